mkdir -p gpurun_out/san
for tool in memcheck racecheck synccheck; do
  log=gpurun_out/san/r2b_sanitizer_$tool.log
  : > $log
  timeout 400 compute-sanitizer --tool $tool python __graft_entry__.py smoke >> $log 2>&1; echo "rc=$?" >> $log
  timeout 900 compute-sanitizer --tool $tool python -m pytest tests/test_ingest.py tests/test_gpu_parity.py -m gpu -q -x -k "packed or device_fasta or warp_per_fragment or (many_references and (parts-3 or parts-overflow))" >> $log 2>&1; echo "rc=$?" >> $log
  grep -E "SUMMARY|passed|failed|rc=" $log | tail -6
done
