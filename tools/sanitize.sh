#!/bin/bash
# compute-sanitizer memcheck + racecheck over the tests that reach every kernel of the partitioned pipeline at small sizes
# (KATs on the partitioned path, the phase-B sieve variants incl. the hand-back to the compacting variant, skewed / chunked /
# pre-packed feeds).  Usage (GPU box): tools/sanitize.sh <out-prefix>
out=${1:-gpurun_out/sanitizer}
sel='(kats and part) or sieve or (skewed and not full) or chunked or prepacked'
for tool in ${TOOLS:-memcheck racecheck}; do
  log=${out}_${tool}.log
  # racecheck instruments every shared-memory access: it gets the small cases only
  if [ $tool = racecheck ]; then sel='(kats and part and not nopreagg) or (sieve and small) or (geometry_on_skewed and small)'; fi
  echo "# compute-sanitizer --tool $tool python -m pytest tests -q -m gpu -k \"$sel\"" > $log
  ( time timeout 500 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests -q -m gpu -x -k "$sel" ) >> $log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $log | tail -3
done
