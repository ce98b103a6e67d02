# scratch: the command of the last ad-hoc GPU call (scripts/gpu_round.sh is the maintained entry)
(timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6)
