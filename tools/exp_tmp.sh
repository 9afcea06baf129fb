mkdir -p gpurun_out
{
timeout 300 python -m pytest tests/test_train_gpu.py -x -q -m gpu -k "stream or oversized or zipf or stress" 2>&1 | tail -3
for d in 0 2 1; do for lf in 0 6; do
echo "DBG=$d lf=$lf"; YTTM_DBG=$d timeout 120 python tools/probe_scan.py 268435456 8 12 $lf
done; done
} > gpurun_out/exp_vec.log 2>&1
