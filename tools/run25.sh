set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 tools/exchange_check.py > gpurun_out/r2_n8_exchange_check.log 2>&1; echo "rc=$?" >> gpurun_out/r2_n8_exchange_check.log
tail -12 gpurun_out/r2_n8_exchange_check.log
bash tools/run_multi.sh 8
