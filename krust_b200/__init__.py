"""krust_b200 -- B200-native canonical k-mer counting engine behind kmerust's counting API.

Layout: csrc/ (sm_100a kernels + the C ABI of include/kmerust_gpu.h), api.py (host-side mirror of the
reference interface over that ABI), dist.py (hash-sharded multi-GPU counting over torch.distributed),
host/ (C++ host: reference-shaped API + CLI).  Importing never needs a GPU; counting always does.
"""
from .api import (BuilderError, GpuError, GpuKmerCounter, InvalidIndexError, KmeRustError, KmerCounter, KmerIndex,  # noqa: F401
                  KmerLength, KmerLengthError, OutputFormat, SequenceFormat, SequenceParseError, compute_histogram,
                  compute_histogram_packed, count_kmers, count_kmers_from_sequences, count_kmers_sequential,
                  count_kmers_streaming, count_kmers_streaming_packed, count_kmers_with_format,
                  count_kmers_with_quality, histogram_stats, kmix_begin, kmix_finish, load_index, owner_of, parse_fastx, read_records,
                  save_index, unpack_many, unpack_to_string, write_counts)

__version__ = "0.1.0"
