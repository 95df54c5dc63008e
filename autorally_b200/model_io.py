"""Model I/O either side of the hot path: the on-disk and on-wire layouts of the
neural-network dynamics model.

* ``.npz`` as written by the reference's trainer (``ML/utils.py:torch_model_to_npz``, :66-90): keys ``dynamics_W{i}``
  (out x in) and ``dynamics_b{i}``, i from 1, float64; read by ``NeuralNetModel::loadParams`` (PI/neural_net_model.cu:73-106).
* torch ``state_dict`` of ``ML/utils.py:setup_model`` (:16-46): ``nn{i}.weight`` / ``nn{i}.bias``, i from 0.
* the ``/model_updater/model`` message (``autorally_msgs/neuralNetModel``: ``structure`` + per-layer ``weight`` / ``bias``)
  as flattened by ``AutorallyPlant::getModel`` (SRC/autorally_plant.cpp:275-301) for
  ``NeuralNetModel::updateModel(description, data)`` (PI/neural_net_model.cu:152-180): ALL weights, layer by layer,
  row-major, THEN all biases -- unlike the ``[W1|b1|W2|b2|...]`` interleave of ``paramsToDevice`` (:120-141).

No GPU and no oracle here; pure numpy.
"""
from __future__ import annotations

import numpy as np

from .params import pack_nn_params, unpack_nn_params


def state_dict_to_npz_dict(state_dict) -> dict:
    """``{'nn0.weight': W, 'nn0.bias': b, ...}`` (torch tensors or arrays) -> ``{'dynamics_W1': W, 'dynamics_b1': b, ...}`` float64."""
    out, wi, bi = {}, 1, 1
    for name, value in state_dict.items():
        arr = np.asarray(value.detach().cpu().numpy() if hasattr(value, "detach") else value, np.float64)
        if "weight" in name:
            out["dynamics_W%d" % wi] = arr
            wi += 1
        elif "bias" in name:
            out["dynamics_b%d" % bi] = arr
            bi += 1
    return out


def npz_dict_to_state_dict(npz) -> dict:
    """Inverse of :func:`state_dict_to_npz_dict` (float64 arrays keyed like the reference's torch model)."""
    n = sum(1 for k in npz.keys() if k.startswith("dynamics_W"))
    out = {}
    for i in range(1, n + 1):
        out["nn%d.weight" % (i - 1)] = np.asarray(npz["dynamics_W%d" % i], np.float64)
        out["nn%d.bias" % (i - 1)] = np.asarray(npz["dynamics_b%d" % i], np.float64).reshape(-1)
    return out


def npz_dict_to_theta(npz):
    """npz dict -> (theta packed as paramsToDevice does, structure)."""
    sd = npz_dict_to_state_dict(npz)
    n = len(sd) // 2
    ws = [sd["nn%d.weight" % i] for i in range(n)]
    bs = [sd["nn%d.bias" % i] for i in range(n)]
    return pack_nn_params(ws, bs)


def theta_to_npz_dict(theta, structure) -> dict:
    ws, bs = unpack_nn_params(theta, structure)
    out = {}
    for i, (w, b) in enumerate(zip(ws, bs), 1):
        out["dynamics_W%d" % i] = np.asarray(w, np.float64)
        out["dynamics_b%d" % i] = np.asarray(b, np.float64)
    return out


def flatten_for_update_model(theta, structure):
    """(description, data) for ``updateModel``: the message layout, all weights then all biases."""
    ws, bs = unpack_nn_params(theta, structure)
    data = np.concatenate([np.asarray(w, np.float32).reshape(-1) for w in ws] + [np.asarray(b, np.float32).reshape(-1) for b in bs])
    return np.asarray(structure, np.int32), data


def theta_from_update_model(description, data):
    """Inverse of :func:`flatten_for_update_model`: back to the ``[W1|b1|W2|b2|...]`` packing."""
    description = [int(v) for v in description]
    data = np.asarray(data, np.float32).reshape(-1)
    ws, off = [], 0
    for nin, nout in zip(description[:-1], description[1:]):
        ws.append(data[off:off + nin * nout].reshape(nout, nin))
        off += nin * nout
    bs = []
    for nout in description[1:]:
        bs.append(data[off:off + nout])
        off += nout
    if off != data.size:
        raise ValueError("update-model payload has %d floats, structure needs %d" % (data.size, off))
    return pack_nn_params(ws, bs)[0]
