mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nmfk_gpu.py -q --deselect tests/test_nmfk_gpu.py::test_nmfk_end_to_end_matches_reference > gpurun_out/t_nmfk.log 2>&1; echo "nmfk rc=$?"
tail -30 gpurun_out/t_nmfk.log
timeout 600 python -m pytest tests/test_nmfk_gpu.py -q -k end_to_end > gpurun_out/t_nmfk_e2e.log 2>&1; echo "e2e rc=$?"
tail -30 gpurun_out/t_nmfk_e2e.log
timeout 200 python -m pytest tests/test_parity_gpu.py -q -x -k "ensemble" > gpurun_out/t_ens.log 2>&1; echo "ens rc=$?"; tail -5 gpurun_out/t_ens.log
