for v in 0 3 6; do echo "== T1_VARIANT=$v"; B200SA_T1_VARIANT=$v python tools/msd_probe.py 3e9 4 msd 2>&1 | grep -E "msd_part1|wall" | cut -c1-100; done
for v in 0 1; do echo "== P_VARIANT=$v"; B200SA_P_VARIANT=$v python tools/msd_probe.py 3e9 4 msd 2>&1 | grep -E "msd_part |wall" | cut -c1-100; done
python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
