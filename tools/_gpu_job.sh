cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "hi_only or 262144" 2>&1 | grep -E "q_grad|hi-only|passed|failed|Error|assert" | tail -12
