// Empty stand-in: PI/param_getter.h includes this header but the MPPI hot path uses nothing from it (oracle/refbuild.py).
