mkdir -p gpurun_out
rm -f gpurun_out/s4g_suite.log
for i in 1 2 3 4 5; do
( timeout 1200 python -m pytest tests -m gpu -q --timeout 900 2>&1 | grep -E "^E  |passed|failed|^FAILED" | cut -c1-900 ) >> gpurun_out/s4g_suite.log 2>&1
echo "---- run $i" >> gpurun_out/s4g_suite.log
done
echo done
