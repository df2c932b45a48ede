set -x
nvidia-smi --query-gpu=name,memory.total --format=csv
python -m pytest tests/test_gpu_engine_golden.py -x -q 2>&1 | tail -25
