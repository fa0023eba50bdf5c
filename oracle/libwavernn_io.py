"""libwavernn `.bin` wire format (reference: vocoder/libwavernn/convert.py:14-59 header + layer order,
:61-84 compress/writeCompressed, :87-175 per-layer savers, :302-352 save order).  TEST INFRASTRUCTURE ONLY:
writes the file the reference's exporter would write for a state_dict so the C++ comparator
(oracle/libwavernn_port.cpp) reads the real format."""
import struct

import numpy as np

from .wavernn_oracle import compress
from .weights import PAD, RES_BLOCKS, UPSAMPLE

ENUM = {"Conv1d": 1, "Conv2d": 2, "BatchNorm1d": 3, "Linear": 4, "GRU": 5, "Stretch2d": 6}


def _hdr(f, kind, name):
    f.write(struct.pack("@i64s", ENUM[kind], name.encode()[:63]))            # convert.py:170-175


def _compressed(f, W):
    w, idx = compress(np.ascontiguousarray(W, np.float32))                   # convert.py:78-84
    f.write(struct.pack("@i", w.size)); f.write(w.tobytes())
    f.write(struct.pack("@i", idx.size)); f.write(idx.tobytes())


def _conv1d(f, sd, p, bias):
    W = sd[p + ".weight"]
    o, i, k = W.shape
    _hdr(f, "Conv1d", p)
    f.write(struct.pack("@iiiii", 4, int(bias), i, o, k)); f.write(np.ascontiguousarray(W, np.float32).tobytes())
    if bias:
        f.write(np.ascontiguousarray(sd[p + ".bias"], np.float32).tobytes())


def _bn(f, sd, p):
    _hdr(f, "BatchNorm1d", p)
    f.write(struct.pack("@iif", 4, sd[p + ".weight"].shape[0], 1e-5))
    for part in (".weight", ".bias", ".running_mean", ".running_var"):       # convert.py:124-133
        f.write(np.ascontiguousarray(sd[p + part], np.float32).tobytes())


def _linear(f, sd, p):
    W = sd[p + ".weight"]
    _hdr(f, "Linear", p)
    f.write(struct.pack("@iii", 4, W.shape[0], W.shape[1]))
    _compressed(f, W)
    f.write(np.ascontiguousarray(sd[p + ".bias"], np.float32).tobytes())


def _gru(f, sd, p):
    wi, wh = sd[p + ".weight_ih_l0"], sd[p + ".weight_hh_l0"]
    bi, bh = sd[p + ".bias_ih_l0"], sd[p + ".bias_hh_l0"]
    hidden, inp = wi.shape[0] // 3, wi.shape[1]
    _hdr(f, "GRU", p)
    f.write(struct.pack("@iii", 4, hidden, inp))
    for W in list(np.vsplit(wi, 3)) + list(np.vsplit(wh, 3)):               # W_ir, W_iz, W_in, W_hr, W_hz, W_hn
        _compressed(f, W)
    for b in list(np.split(bi, 3)) + list(np.split(bh, 3)):
        f.write(np.ascontiguousarray(b, np.float32).tobytes())


def write_bin(path, sd):
    with open(path, "wb") as f:
        f.write(struct.pack("@iiii", RES_BLOCKS, len(UPSAMPLE), int(np.prod(UPSAMPLE)), PAD))   # convert.py:55
        r = "upsample.resnet"
        _conv1d(f, sd, r + ".conv_in", False)
        _bn(f, sd, r + ".batch_norm")
        for i in range(RES_BLOCKS):
            p = "%s.layers.%d" % (r, i)
            _conv1d(f, sd, p + ".conv1", False); _bn(f, sd, p + ".batch_norm1")
            _conv1d(f, sd, p + ".conv2", False); _bn(f, sd, p + ".batch_norm2")
        _conv1d(f, sd, r + ".conv_out", True)
        _hdr(f, "Stretch2d", "resnet_stretch"); f.write(struct.pack("@ii", int(np.prod(UPSAMPLE)), 1))
        for idx, s in zip((1, 3, 5), UPSAMPLE):
            _hdr(f, "Stretch2d", "stretch%d" % idx); f.write(struct.pack("@ii", s, 1))
            w = np.ascontiguousarray(sd["upsample.up_layers.%d.weight" % idx], np.float32).reshape(-1)
            _hdr(f, "Conv2d", "up%d" % idx); f.write(struct.pack("@ii", 4, w.size)); f.write(w.tobytes())
        _linear(f, sd, "I"); _gru(f, sd, "rnn1"); _gru(f, sd, "rnn2")
        _linear(f, sd, "fc1"); _linear(f, sd, "fc2"); _linear(f, sd, "fc3")
