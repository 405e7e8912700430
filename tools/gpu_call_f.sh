set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
python tools/sanitize_smoke.py 2>&1 | tail -3
compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_smoke.py > gpurun_out/sanitizer_memcheck_r1.log 2>&1; tail -6 gpurun_out/sanitizer_memcheck_r1.log
compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_smoke.py > gpurun_out/sanitizer_racecheck_r1.log 2>&1; tail -6 gpurun_out/sanitizer_racecheck_r1.log
compute-sanitizer --tool synccheck --print-limit 20 python tools/sanitize_smoke.py > gpurun_out/sanitizer_synccheck_r1.log 2>&1; tail -4 gpurun_out/sanitizer_synccheck_r1.log
