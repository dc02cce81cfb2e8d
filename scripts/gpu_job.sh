timeout 50 python -m pytest tests/test_shard.py -m gpu -x -q -k "0-nccl or 1-ce" 2>&1 | tail -3
