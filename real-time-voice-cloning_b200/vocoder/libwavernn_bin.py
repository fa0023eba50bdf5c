"""Reader of libwavernn `.bin` checkpoints (the file vocoder_convert_model.py exports for the C++ vocoder), so that
`load_model(path, voc_type='libwavernn')` runs the exported -- usually pruned -- fatchord model on the B200 engine.

Wire format (reference: vocoder/libwavernn/convert.py:55 file header `@iiii` = res_blocks, #upsample layers, total scale,
pad; :170-175 layer header `@i64s` = type enum 1..6 + printed layer name; payloads: Linear :87-96 `@iii` elSize, rows,
cols + compressed matrix + bias; GRU :135-162 `@iii` elSize, hidden, input + six compressed matrices W_ir, W_iz, W_in,
W_hr, W_hz, W_hn + six biases; Conv1d :98-108 `@iiiii` elSize, has_bias, in, out, kernel; Conv2d :110-119 `@ii` elSize,
kernel; BatchNorm1d :121-133 `@iif` elSize, features, eps + weight, bias, running_mean, running_var; Stretch2d :164-167
`@ii` x_scale, y_scale.  Compressed matrix :61-84 / wavernn.h:23-92: `int nW, float[nW]` = the kept 1x4 column groups in
row-major order, `int nIdx, uint8[nIdx]` = group-column indices per row, 255 ends a row (rows + 1 markers).  Layer order
:57-59, 302-352: resnet, upsample, I, rnn1, rnn2, fc1, fc2, fc3.)

The decoder goes by layer type and order, not by the printed names.  Column indices are decoded as unsigned bytes
(the reference's C++ reader has bug Q12 for matrices with more than 127 groups per row; the file itself is fine).
"""
import struct

import numpy as np

CONV1D, CONV2D, BATCHNORM1D, LINEAR, GRU, STRETCH2D = 1, 2, 3, 4, 5, 6
SPARSE_GROUP = 4     # hparams.sparse_group (config/hparams.py:270)


class BinFormatError(ValueError):
    pass


class _Reader:
    def __init__(self, buf):
        self.buf, self.pos = buf, 0

    def unpack(self, fmt):
        n = struct.calcsize(fmt)
        if self.pos + n > len(self.buf):
            raise BinFormatError("truncated libwavernn file (wanted %d bytes at offset %d)" % (n, self.pos))
        v = struct.unpack_from(fmt, self.buf, self.pos)
        self.pos += n
        return v

    def array(self, dtype, count):
        n = np.dtype(dtype).itemsize * count
        if count < 0 or self.pos + n > len(self.buf):
            raise BinFormatError("truncated libwavernn file (wanted %d bytes at offset %d)" % (n, self.pos))
        a = np.frombuffer(self.buf, dtype=dtype, count=count, offset=self.pos).copy()
        self.pos += n
        return a

    def header(self, want):
        kind, _name = self.unpack("@i64s")
        if kind != want:
            raise BinFormatError("layer type %d where %d was expected (offset %d): not a fatchord libwavernn export" % (kind, want, self.pos))

    def compressed(self, rows, cols):
        """wavernn.h:23-92 / convert.py:61-84 -> dense (rows, cols) float32."""
        (nw,) = self.unpack("@i")
        w = self.array(np.float32, nw)
        (nidx,) = self.unpack("@i")
        idx = self.array(np.uint8, nidx)
        W = np.zeros((rows, cols), np.float32)
        ends = np.flatnonzero(idx == 255)
        if ends.size < rows or nw % SPARSE_GROUP:
            raise BinFormatError("compressed matrix does not describe %d rows" % rows)
        start, k = 0, 0
        for r in range(rows):
            for c in idx[start:ends[r]]:
                if (int(c) + 1) * SPARSE_GROUP > cols or k + SPARSE_GROUP > nw:
                    raise BinFormatError("compressed matrix index out of range")
                W[r, int(c) * SPARSE_GROUP:(int(c) + 1) * SPARSE_GROUP] = w[k:k + SPARSE_GROUP]
                k += SPARSE_GROUP
            start = ends[r] + 1
        if k != nw:
            raise BinFormatError("compressed matrix has %d stray weights" % (nw - k))
        return W

    def conv1d(self, sd, name):
        self.header(CONV1D)
        el, has_bias, cin, cout, k = self.unpack("@iiiii")
        _check_el(el)
        sd[name + ".weight"] = self.array(np.float32, cout * cin * k).reshape(cout, cin, k)
        if has_bias:
            sd[name + ".bias"] = self.array(np.float32, cout)

    def batchnorm(self, sd, name):
        self.header(BATCHNORM1D)
        el, n, _eps = self.unpack("@iif")
        _check_el(el)
        for part in (".weight", ".bias", ".running_mean", ".running_var"):
            sd[name + part] = self.array(np.float32, n)

    def linear(self, sd, name):
        self.header(LINEAR)
        el, rows, cols = self.unpack("@iii")
        _check_el(el)
        sd[name + ".weight"] = self.compressed(rows, cols)
        sd[name + ".bias"] = self.array(np.float32, rows)

    def gru(self, sd, name):
        self.header(GRU)
        el, hidden, inp = self.unpack("@iii")
        _check_el(el)
        wi = [self.compressed(hidden, inp) for _ in range(3)]
        wh = [self.compressed(hidden, hidden) for _ in range(3)]
        b = [self.array(np.float32, hidden) for _ in range(6)]
        sd[name + ".weight_ih_l0"] = np.vstack(wi)
        sd[name + ".weight_hh_l0"] = np.vstack(wh)
        sd[name + ".bias_ih_l0"] = np.concatenate(b[:3])
        sd[name + ".bias_hh_l0"] = np.concatenate(b[3:])

    def stretch(self):
        self.header(STRETCH2D)
        return self.unpack("@ii")


def _check_el(el):
    if el != 4:
        raise BinFormatError("element size %d: only float32 exports exist (convert.py:12)" % el)


def read_bin(path):
    """-> (state_dict of numpy arrays under the fatchord WaveRNN names, meta dict).  meta: res_blocks, upsample_factors, pad,
    n_classes (fc3 rows: 2**bits for RAW, 30 for MOL)."""
    with open(path, "rb") as f:
        buf = f.read()
    r = _Reader(buf)
    res_blocks, n_up, total, pad = r.unpack("@iiii")
    if not (0 < res_blocks <= 64 and 0 < n_up <= 8):
        raise BinFormatError("implausible libwavernn header (res_blocks=%d, upsample layers=%d)" % (res_blocks, n_up))
    sd = {}
    rn = "upsample.resnet"
    r.conv1d(sd, rn + ".conv_in")
    r.batchnorm(sd, rn + ".batch_norm")
    for i in range(res_blocks):
        p = "%s.layers.%d" % (rn, i)
        r.conv1d(sd, p + ".conv1"); r.batchnorm(sd, p + ".batch_norm1")
        r.conv1d(sd, p + ".conv2"); r.batchnorm(sd, p + ".batch_norm2")
    r.conv1d(sd, rn + ".conv_out")
    (scale, _y) = r.stretch()
    if scale != total:
        raise BinFormatError("resnet stretch %d != total scale %d" % (scale, total))
    factors = []
    for i in range(n_up):
        (s, _y) = r.stretch()
        factors.append(int(s))
        r.header(CONV2D)
        el, k = r.unpack("@ii")
        _check_el(el)
        sd["upsample.up_layers.%d.weight" % (2 * i + 1)] = r.array(np.float32, k).reshape(1, 1, 1, k)
    if int(np.prod(factors)) != total:
        raise BinFormatError("upsample factors %s do not multiply to %d" % (factors, total))
    r.linear(sd, "I")
    r.gru(sd, "rnn1")
    r.gru(sd, "rnn2")
    r.linear(sd, "fc1")
    r.linear(sd, "fc2")
    r.linear(sd, "fc3")
    if r.pos != len(buf):
        raise BinFormatError("%d trailing bytes: not a fatchord libwavernn export" % (len(buf) - r.pos))
    sd["step"] = np.zeros(1, np.int64)            # the export drops the training step
    meta = dict(res_blocks=int(res_blocks), upsample_factors=tuple(factors), pad=int(pad), n_classes=int(sd["fc3.weight"].shape[0]))
    return sd, meta
