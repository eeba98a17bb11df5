timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "scan_front_end or small_scan_retry or end_to_end_trajectory" > gpurun_out/r02b_memcheck.log 2>&1; echo memcheck rc=$?; tail -6 gpurun_out/r02b_memcheck.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
