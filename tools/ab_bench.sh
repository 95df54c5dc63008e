#!/bin/bash
# A/B on ONE box: the library of the last commit (lib/exp/libmppi_b200_head.so) against the working tree's, alternating.
L=autorally_b200/lib
cp $L/libmppi_b200.so /tmp/new.so
for rep in 1 2 3; do for which in head new; do
  if [ $which = head ]; then cp $L/exp/libmppi_b200_head.so $L/libmppi_b200.so; else cp /tmp/new.so $L/libmppi_b200.so; fi
  echo -n "$which: "; python bench.py --no-large --no-cpu --steps 400 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['ms_per_step_warm_l2'], d['e2e']['p50_ms'], d['roofline']['kernel_ms'])"
done; done
cp /tmp/new.so $L/libmppi_b200.so
