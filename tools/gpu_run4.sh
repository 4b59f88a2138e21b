mkdir -p gpurun_out
timeout 300 python tools/diag_tf32_accuracy.py > gpurun_out/r2_diag_tf32_b.log 2>&1
timeout 600 python -m pytest tests/test_tf32_gpu.py -q --timeout 300 -s 2>&1 | grep "rel err\|passed\|failed\|Error" > gpurun_out/r2_t_tf32_b.log
timeout 900 python -m pytest tests/test_configs_gpu.py::test_c1_fgsm_256_style_fusion_vs_oracle "tests/test_engine_gpu.py::test_fp32_mode_meets_north_star_tolerance" tests/test_fullsize_gpu.py tests/test_surface_gpu.py::test_cal_result_metrics_at_full_resolution -q --timeout 600 -rA -s > gpurun_out/r2_t4.log 2>&1
echo done
