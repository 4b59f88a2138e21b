mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_configs_gpu.py::test_c1_fgsm_256_style_fusion_vs_oracle "tests/test_engine_gpu.py::test_fp32_mode_meets_north_star_tolerance" tests/test_fullsize_gpu.py tests/test_surface_gpu.py::test_cal_result_metrics_at_full_resolution -q --timeout 600 -rA -s > gpurun_out/r2_t2.log 2>&1
timeout 300 python tools/diag_patch512.py 512 2e3 > gpurun_out/r2_diag_patch.log 2>&1
timeout 300 python tools/diag_patch512.py 512 1.0 >> gpurun_out/r2_diag_patch.log 2>&1
echo done
