set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=8
P=29700
run() { name=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N "$@" > gpurun_out/r2_n${N}_$name.json 2> gpurun_out/r2_n${N}_$name.err; P=$((P+1)); tail -c 300 gpurun_out/r2_n${N}_$name.json; tail -2 gpurun_out/r2_n${N}_$name.err; }
run c4g --workload c4 --steps 5 --warmup 3 --no-cpu
run c3 --steps 20 --warmup 3 --no-e2e --no-cpu
