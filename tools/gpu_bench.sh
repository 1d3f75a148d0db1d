set -x
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -15
python bench.py --steps 30 --warmup 3 --solver 0 --cpu-sample 8 2>&1 | tail -1
python bench.py --steps 30 --warmup 3 --solver 0 --batch 16384 --cpu-sample 8 2>&1 | tail -1
