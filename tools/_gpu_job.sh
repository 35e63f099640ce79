cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
bash tools/verify_round.sh 2>&1 | tail -12
bash tools/run_configs.sh 2>&1 | tail -16
