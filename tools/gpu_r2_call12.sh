set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/ao_mo_once.py > gpurun_out/r2_ao_mo.txt 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ws_dgemm -s 8 -c 2 -o gpurun_out/r2_ao_mo_gemm python tools/ao_mo_once.py > gpurun_out/ncu_ao.log 2>&1
cat gpurun_out/r2_ao_mo.txt
