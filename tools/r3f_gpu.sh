for e in "SP2_TAIL_PIPE=0" "SP2_MID_PIPE=0" "SP2_NO_PERSIST=1" "SP2_KECCAK_THREAD=1"; do
  echo "== $e"; env $e timeout 900 python -m pytest tests/test_gpu_sumcheck.py tests/test_gpu_spartan.py -m gpu -x -q 2>&1 | tail -2
done
