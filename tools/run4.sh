set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -x -q > gpurun_out/r2_tc1.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tc1.log
tail -40 gpurun_out/r2_tc1.log
