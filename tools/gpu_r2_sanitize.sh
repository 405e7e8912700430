# one compute-sanitizer tool per gpurun call (B200_PROFILING.md); usage: bash tools/gpu_r2_sanitize.sh memcheck|racecheck
set -x
TOOL=${1:-memcheck}
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_plain.log 2>&1 && \
timeout 1500 compute-sanitizer --tool $TOOL --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitizer_${TOOL}_smoke.log 2>&1
tail -15 gpurun_out/r2_sanitizer_${TOOL}_smoke.log
