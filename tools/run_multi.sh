# usage: bash tools/run_multi.sh N   (inside gpurun --gpus N): the multi-GPU bench lines of this round -> gpurun_out/r2_nN_*.json
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$1
P=29600
run() { name=$1; shift; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N "$@" > gpurun_out/r2_n${N}_$name.json 2> gpurun_out/r2_n${N}_$name.err; P=$((P+1)); tail -c 400 gpurun_out/r2_n${N}_$name.json; tail -2 gpurun_out/r2_n${N}_$name.err; }
run c4g --workload c4 --steps 5 --warmup 3
run c5 --workload c5 --steps 10 --warmup 3
run c3_strong --scaling strong --steps 30 --warmup 3 --no-e2e --no-cpu
run c3 --steps 20 --warmup 3 --no-e2e --no-cpu
if [ "$2" = "full" ]; then run c4u --workload c4 --ungrouped --steps 5 --warmup 3; fi
