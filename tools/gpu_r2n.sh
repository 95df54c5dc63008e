#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -q -rf 2>&1 | tail -12
python bench.py --steps 20 --warmup 5 --no-large --no-cpu | cut -c1-1400
MPPI_NO_PIPELINED_SAMPLER=1 python bench.py --steps 20 --warmup 5 --no-large --no-cpu | cut -c1-1400
