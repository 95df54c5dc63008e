"""CPU: model I/O formats either side of the hot path."""
import os
import sys
import types

import numpy as np
import pytest

from autorally_b200 import model_io
from autorally_b200.params import load_nn_npz, unpack_nn_params

REF_ML = "/root/reference/autorally_control/src/path_integral/scripts/ml_pipeline"


def test_npz_state_dict_theta_round_trips(tmp_path, models):
    theta, structure = models["autorally_nnet_theta"], models["autorally_nnet_structure"]
    npz = model_io.theta_to_npz_dict(theta, structure)
    assert sorted(npz) == ["dynamics_W1", "dynamics_W2", "dynamics_W3", "dynamics_b1", "dynamics_b2", "dynamics_b3"]
    assert npz["dynamics_W1"].shape == (32, 6) and npz["dynamics_W1"].dtype == np.float64   # out x in, float64 on disk
    sd = model_io.npz_dict_to_state_dict(npz)
    assert list(sd) == ["nn0.weight", "nn0.bias", "nn1.weight", "nn1.bias", "nn2.weight", "nn2.bias"]
    back = model_io.state_dict_to_npz_dict(sd)
    for k in npz:
        np.testing.assert_array_equal(back[k], npz[k])
    th2, st2 = model_io.npz_dict_to_theta(back)
    np.testing.assert_array_equal(th2, theta)
    np.testing.assert_array_equal(st2, structure)
    np.savez(tmp_path / "m.npz", **npz)
    th3, st3 = load_nn_npz(tmp_path / "m.npz")   # the loader the C++ NeuralNetModel::loadParams mirrors
    np.testing.assert_array_equal(th3, theta)


def test_update_model_message_layout(models):
    """All weights (layer by layer, row-major) then all biases: AutorallyPlant::getModel (SRC/autorally_plant.cpp:275-301)."""
    theta, structure = models["wider_deeper_theta"], models["wider_deeper_structure"]
    description, data = model_io.flatten_for_update_model(theta, structure)
    ws, bs = unpack_nn_params(theta, structure)
    nw = sum(w.size for w in ws)
    assert data.size == theta.size and list(description) == list(structure)
    np.testing.assert_array_equal(data[:ws[0].size], ws[0].reshape(-1))
    np.testing.assert_array_equal(data[nw:nw + bs[0].size], bs[0])
    np.testing.assert_array_equal(data[-bs[-1].size:], bs[-1])
    assert not np.array_equal(data, theta)   # it is NOT the [W1|b1|W2|b2|...] interleave of paramsToDevice
    np.testing.assert_array_equal(model_io.theta_from_update_model(description, data), theta)
    with pytest.raises(ValueError):
        model_io.theta_from_update_model(description, data[:-1])


@pytest.mark.filterwarnings("ignore::SyntaxWarning")
@pytest.mark.skipif(not os.path.isdir(REF_ML), reason="reference checkout not present (build container only)")
def test_against_the_reference_trainer_functions(tmp_path, models):
    """ML/utils.py torch_model_to_npz / npz_to_torch_model, imported unmodified (matplotlib is only used for plots)."""
    torch = pytest.importorskip("torch")
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, REF_ML)
    try:
        import utils as ref_utils
    finally:
        sys.path.remove(REF_ML)
    theta, structure = models["autorally_nnet_theta"], models["autorally_nnet_structure"]
    np.savez(tmp_path / "in.npz", **model_io.theta_to_npz_dict(theta, structure))
    model = ref_utils.setup_model(layers=[int(v) for v in structure])
    model = ref_utils.npz_to_torch_model(str(tmp_path / "in.npz"), model)
    # our state_dict view of the npz equals what the reference loaded into its torch model
    sd = model_io.npz_dict_to_state_dict(np.load(tmp_path / "in.npz"))
    for k, v in model.state_dict().items():
        np.testing.assert_array_equal(v.numpy(), sd[k].reshape(v.shape))
    # and the reference's writer produces the npz our reader maps back to the same packed parameters
    ref_utils.torch_model_to_npz(model, str(tmp_path))
    th, st = load_nn_npz(tmp_path / "model.npz")
    np.testing.assert_array_equal(th, theta)
    ours = model_io.state_dict_to_npz_dict(model.state_dict())
    z = np.load(tmp_path / "model.npz")
    for k in z.files:
        np.testing.assert_array_equal(ours[k].reshape(z[k].shape), z[k])
