set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests9.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests9.log
tail -25 gpurun_out/r2_tests9.log
P=29551
run2() { name=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 "$@" > gpurun_out/r2_n2_$name.json 2> gpurun_out/r2_n2_$name.err; P=$((P+1)); tail -c 600 gpurun_out/r2_n2_$name.json; tail -2 gpurun_out/r2_n2_$name.err; }
run2 c3 --steps 10 --warmup 3 --no-cpu
run2 c3_strong --scaling strong --steps 20 --warmup 3 --no-e2e --no-cpu
run2 c4g --workload c4 --steps 5 --warmup 3
run2 c4u --workload c4 --ungrouped --steps 5 --warmup 3
run2 c5 --workload c5 --steps 5 --warmup 3
