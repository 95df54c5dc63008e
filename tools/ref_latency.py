import numpy as np, sys, os
sys.path.insert(0, os.getcwd())
from oracle import reference as ref
from autorally_b200.params import make_ellipse_costmap
from tests.common import cost_params_for, straight_controls, top_state
models = np.load("tests/golden/ref_models.npz")
costmap = make_ellipse_costmap(); cp = cost_params_for(costmap)
state, U = top_state(4.0), straight_controls(100)
for kind, tag, neg in ((ref.REF_NN64_1920, "wider_deeper", False), (ref.REF_NN_1920, "autorally_nnet", True)):
    with ref.ReferenceController(kind, models[tag + "_theta"], costmap, cp, negate_yaw_der=neg) as rc:
        rc.set_controls(U, np.zeros(4, np.float32))
        print(tag, "reference computeControl ms:", rc.time_compute_control(state, reps=30))
