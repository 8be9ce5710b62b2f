python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -2
python tools/msd_probe.py 3e9 4 msd 2>&1 | grep -E "wall|local_sort" | cut -c1-110
