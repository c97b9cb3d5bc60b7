//! Raw bindings of include/kmerust_gpu.h (ABI v1).  NOT compiled in the build image (no Rust
//! toolchain there); kept mechanically in step with the header -- same order, same field layout
//! (tests/test_host_cpu.py pins the struct sizes 48 / 88 / 56 that these #[repr(C)] types must have).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const KMG_ABI_VERSION: u32 = 1;

pub type kmg_status = c_int;
pub const KMG_OK: kmg_status = 0;
pub const KMG_ERR_INVALID_K: kmg_status = 1;
pub const KMG_ERR_INVALID_ARG: kmg_status = 2;
pub const KMG_ERR_CUDA: kmg_status = 3;
pub const KMG_ERR_OOM: kmg_status = 4;
pub const KMG_ERR_TABLE_FULL: kmg_status = 5;
pub const KMG_ERR_STATE: kmg_status = 6;
pub const KMG_ERR_IO: kmg_status = 7;
pub const KMG_ERR_ABI: kmg_status = 8;
pub const KMG_ERR_CAPACITY: kmg_status = 9;
pub const KMG_ERR_PARSE: kmg_status = 10;

pub const KMG_FLAG_FORCE_HASH: u32 = 1;
pub const KMG_FLAG_FORCE_DIRECT: u32 = 2;
pub const KMG_FLAG_NO_PREAGG: u32 = 4;
pub const KMG_FLAG_FORCE_PARTITIONED: u32 = 8;

#[repr(C)]
#[derive(Debug, Clone, Copy)]
pub struct kmg_config {
    pub abi_version: u32,
    pub k: u32,
    pub device: i32,
    pub flags: u32,
    pub has_min_quality: u8,
    pub min_quality: u8,
    pub parts_log2: u8,
    pub reserved: [u8; 5],
    pub expected_distinct: u64,
    pub batch_bases: u64,
    pub stream: *mut c_void,
}

#[repr(C)]
#[derive(Debug, Clone, Copy, Default)]
pub struct kmg_summary {
    pub n_records: u64,
    pub n_bases: u64,
    pub n_windows: u64,
    pub n_distinct: u64,
    pub max_count: u64,
    pub table_capacity: u64,
    pub path: u32,
    pub n_grows: u32,
    pub kernel_ns: u64,
    pub h2d_bytes: u64,
    pub scan_ns: u64,
    pub consolidate_ns: u64,
}

#[repr(C)]
#[derive(Debug, Clone, Copy)]
pub struct kmg_batch {
    pub bases2bit: *mut u64,
    pub valid_bits: *mut u32,
    pub start_bits: *mut u32,
    pub capacity_bases: u64,
    pub n_bases: u64,
    pub n_records: u64,
    pub slot: u32,
}

#[repr(C)]
pub struct kmg_ctx {
    _private: [u8; 0],
}

extern "C" {
    pub fn kmg_abi_version() -> u32;
    pub fn kmg_status_string(s: kmg_status) -> *const c_char;
    pub fn kmg_last_error(ctx: *const kmg_ctx) -> *const c_char;
    pub fn kmg_create(cfg: *const kmg_config, out: *mut *mut kmg_ctx) -> kmg_status;
    pub fn kmg_destroy(ctx: *mut kmg_ctx);
    pub fn kmg_reset(ctx: *mut kmg_ctx) -> kmg_status;
    pub fn kmg_count_ascii(ctx: *mut kmg_ctx, seq: *const u8, qual: *const u8, offsets: *const u64, n_records: u64) -> kmg_status;
    pub fn kmg_acquire_batch(ctx: *mut kmg_ctx, batch: *mut kmg_batch) -> kmg_status;
    pub fn kmg_submit_batch(ctx: *mut kmg_ctx, batch: *const kmg_batch) -> kmg_status;
    pub fn kmg_count_ascii_device(ctx: *mut kmg_ctx, d_seq: *const u8, d_qual: *const u8, d_offsets: *const u64, n_records: u64, n_bytes: u64) -> kmg_status;
    pub fn kmg_insert_keys_device(ctx: *mut kmg_ctx, d_keys: *const u64, d_counts: *const u64, n: u64) -> kmg_status;
    pub fn kmg_extract_keys_device(ctx: *mut kmg_ctx, d_seq: *const u8, d_qual: *const u8, d_offsets: *const u64, n_records: u64, n_bytes: u64,
                                   n_shards: u32, d_keys_out: *mut u64, cap: u64, shard_counts_out: *mut u64) -> kmg_status;
    pub fn kmg_owner_of(canonical_key: u64, n_shards: u32) -> u32;
    pub fn kmg_finalize(ctx: *mut kmg_ctx, summary: *mut kmg_summary) -> kmg_status;
    pub fn kmg_export_counts(ctx: *mut kmg_ctx, min_count: u64, sorted: c_int, keys: *mut u64, counts: *mut u64, cap: u64, n_out: *mut u64) -> kmg_status;
    pub fn kmg_export_counts_device(ctx: *mut kmg_ctx, min_count: u64, sorted: c_int, d_keys: *mut u64, d_counts: *mut u64, cap: u64, n_out: *mut u64) -> kmg_status;
    pub fn kmg_histogram(ctx: *mut kmg_ctx, min_count: u64, count_vals: *mut u64, freqs: *mut u64, cap: u64, n_out: *mut u64) -> kmg_status;
    pub fn kmg_save_kmix(ctx: *mut kmg_ctx, path: *const c_char) -> kmg_status;
    pub fn kmg_progress(ctx: *const kmg_ctx, records: *mut u64, bases: *mut u64) -> kmg_status;
    pub fn kmg_kernel_launches() -> u64;
    pub fn kmg_synth_uniform_device(ctx: *mut kmg_ctx, seed: u64, first_base: u64, n: u64, d_out: *mut u8) -> kmg_status;
    pub fn kmg_parse_fastx(buf: *const u8, len: u64, is_fastq: c_int, seq_out: *mut u8, qual_out: *mut u8, offsets_out: *mut u64,
                           max_records: u64, n_records_out: *mut u64, errbuf: *mut c_char, errbuf_len: usize) -> kmg_status;
}
