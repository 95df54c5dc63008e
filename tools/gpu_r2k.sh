#!/bin/bash
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python tools/profile_step.py --rollouts 1920 --steps 50
python tools/profile_step.py --rollouts 1920 --steps 50 --variant 11
python tools/profile_step.py --rollouts 4096 --steps 50
python bench.py --steps 20 --warmup 5 --no-large --no-cpu | cut -c1-1200
