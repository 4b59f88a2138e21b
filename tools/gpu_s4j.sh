mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_nway_gpu.py -m gpu -q -s -k "ffhq_5way" --timeout 600 2>&1 | grep -E "^\[|passed|failed|^E  " | cut -c1-500 ) > gpurun_out/s4j_5way.log 2>&1
echo done
