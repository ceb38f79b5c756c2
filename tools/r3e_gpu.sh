timeout 1200 python -m pytest tests/test_gpu_neutronnova.py tests/test_gpu_neutronnova_snark.py -m gpu -x -q 2>&1 | tail -4
python tools/nn_snark_time.py 32 2>&1 | tail -1 | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
