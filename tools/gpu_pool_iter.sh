timeout 600 python tools/check_gemm.py all > gpurun_out/check_gemm.txt 2>&1; echo "check_gemm rc=$?"; grep -c " ok " gpurun_out/check_gemm.txt; grep -i "fail\|error\|Traceback" gpurun_out/check_gemm.txt | head -20
python tools/gemm_ab.py pool 2>&1 | grep -E "^tn|^wg"
python tools/gemm_ab.py wg 2>&1 | grep -E "^tn|^wg"
