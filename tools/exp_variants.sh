for v in 0 7; do B200SA_T1_VARIANT=$v python tools/msd_probe.py 3e9 4 msd 2>&1 | grep -E "wall|part1" | cut -c1-110; done
B200SA_T1_VARIANT=7 python -m pytest tests/test_gpu_parity.py -x -q -k "dna_1M or repeat_rich or dna_100003" 2>&1 | tail -2
