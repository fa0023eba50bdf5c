"""The product-side reader of libwavernn `.bin` exports (vocoder/libwavernn_bin.py) against the
oracle's writer of the reference's format; `load_model(path, voc_type='libwavernn')` on the GPU."""
import numpy as np
import pytest

from oracle import libwavernn_io, weights


def _reader():
    import rtvc_b200  # noqa: F401
    from rtvc_b200.vocoder import libwavernn_bin
    return libwavernn_bin


@pytest.mark.parametrize("mode,bits,prune", [("RAW", 9, None), ("RAW", 9, 0.9), ("RAW", 10, 0.9), ("MOL", 9, 0.9)])
def test_read_bin_round_trip(tmp_path, mode, bits, prune):
    sd = weights.make_state_dict(seed=3, bits=bits, mode=mode)
    if prune:
        sd = weights.prune_state_dict(sd, z=prune)
    path = tmp_path / "m.bin"
    libwavernn_io.write_bin(path, sd)
    got, meta = _reader().read_bin(str(path))
    assert meta == dict(res_blocks=10, upsample_factors=(5, 5, 8), pad=2, n_classes=30 if mode == "MOL" else 2 ** bits)
    for k, v in sd.items():
        if k == "step" or k.endswith("num_batches_tracked"):
            continue                                  # not part of the export (convert.py:121-133)
        assert k in got, k
        np.testing.assert_array_equal(got[k], np.asarray(v, np.float32).reshape(got[k].shape), err_msg=k)
    assert set(got) - set(sd) <= {"step"}


def test_read_bin_rejects_damaged_files(tmp_path):
    rd = _reader()
    sd = weights.make_state_dict(seed=3, bits=9, mode="RAW")
    path = tmp_path / "m.bin"
    libwavernn_io.write_bin(path, sd)
    raw = path.read_bytes()
    (tmp_path / "short.bin").write_bytes(raw[: len(raw) // 2])
    with pytest.raises(rd.BinFormatError):
        rd.read_bin(str(tmp_path / "short.bin"))
    (tmp_path / "long.bin").write_bytes(raw + b"\0" * 8)
    with pytest.raises(rd.BinFormatError):
        rd.read_bin(str(tmp_path / "long.bin"))
    bad = bytearray(raw)
    bad[16] = 9                                       # first layer header: unknown type enum
    (tmp_path / "bad.bin").write_bytes(bytes(bad))
    with pytest.raises(rd.BinFormatError):
        rd.read_bin(str(tmp_path / "bad.bin"))


def test_load_model_libwavernn_missing_file():
    import rtvc_b200  # noqa: F401
    from rtvc_b200.vocoder import inference
    with pytest.raises(RuntimeError, match="Cannot open file."):      # WaveRNNVocoder.cpp:24-26
        inference.load_model("/nonexistent/model.bin", voc_type="libwavernn", verbose=False)


@pytest.mark.gpu
def test_load_model_libwavernn_matches_state_dict(tmp_path):
    """The exported pruned model vocodes exactly like the state_dict it was exported from (block-sparse loop)."""
    import rtvc_b200  # noqa: F401
    from rtvc_b200.vocoder import inference
    sd = weights.prune_state_dict(weights.make_state_dict(seed=11, bits=9, mode="RAW"), z=0.9)
    path = tmp_path / "m.bin"
    libwavernn_io.write_bin(path, sd)
    mel = weights.synthetic_mel(40, seed=5)
    import copy
    from rtvc_b200.config.hparams import wavernn_fatchord
    hp = copy.deepcopy(wavernn_fatchord)
    hp.bits, hp.mode = 9, "RAW"
    inference.load_state(sd, override_hp_fatchord=hp)
    inference.set_seed(7)
    a = inference.infer_waveform(mel, target=1200, overlap=200)
    inference.load_model(str(path), voc_type="libwavernn", verbose=False)
    assert inference.is_loaded()
    inference.set_seed(7)
    b = inference.infer_waveform(mel, target=1200, overlap=200)
    assert a.shape == b.shape == ((40 - 1) * 200,) and a.dtype == np.float64
    np.testing.assert_array_equal(a, b)


def _c_loader(path, device=0):
    import ctypes as C
    import rtvc_b200  # noqa: F401
    from rtvc_b200 import _native
    lib = _native.load()
    h = C.c_void_p()
    err = C.create_string_buffer(256)
    rc = lib.wrnn_create_from_bin(str(path).encode(), device, C.byref(h), err, 256)
    return lib, rc, h, err.value.decode()


def test_c_level_bin_loader_errors(tmp_path):
    """wrnn_create_from_bin (the C ABI's WaveRNNVocoder::loadWeights): errors are decided before any GPU is touched."""
    from rtvc_b200 import _native
    lib, rc, h, msg = _c_loader("/nonexistent/model.bin")
    assert rc == _native.ERR_INVALID and not h.value and msg == "Cannot open file."        # WaveRNNVocoder.cpp:24-26
    sd = weights.make_state_dict(seed=3, bits=9, mode="RAW")
    path = tmp_path / "m.bin"
    libwavernn_io.write_bin(path, sd)
    raw = path.read_bytes()
    (tmp_path / "short.bin").write_bytes(raw[: len(raw) // 2])
    lib, rc, h, msg = _c_loader(tmp_path / "short.bin")
    assert rc == _native.ERR_SHAPE and not h.value and "truncated" in msg
    (tmp_path / "long.bin").write_bytes(raw + b"\0" * 8)
    lib, rc, h, msg = _c_loader(tmp_path / "long.bin")
    assert rc == _native.ERR_SHAPE and "trailing" in msg
    bad = bytearray(raw)
    bad[16] = 9
    (tmp_path / "bad.bin").write_bytes(bytes(bad))
    lib, rc, h, msg = _c_loader(tmp_path / "bad.bin")
    assert rc == _native.ERR_SHAPE and "layer type" in msg


@pytest.mark.gpu
def test_c_level_bin_loader_matches_python_reader(tmp_path):
    """An engine created by wrnn_create_from_bin generates the same samples as one loaded through the Python reader."""
    import ctypes as C
    from tests.util import norm_mel
    import rtvc_b200  # noqa: F401
    from rtvc_b200 import _native
    from rtvc_b200.vocoder import inference
    sd = weights.prune_state_dict(weights.make_state_dict(seed=3, bits=9, mode="RAW"), z=0.9)
    path = tmp_path / "m.bin"
    libwavernn_io.write_bin(path, sd)
    lib, rc, h, msg = _c_loader(path)
    assert rc == _native.OK and h.value, msg
    try:
        inference.load_model(str(path), voc_type="libwavernn", verbose=False)
        m = inference._model[0]
        assert abs(lib.wrnn_sparsity(h) - m.sparsity) < 1e-12 and m.sparsity > 0.8
        mel = norm_mel(30, 2)
        a = m.generate_debug(mel, True, 400, 100, seed=3, max_steps=64, precision=_native.PREC_F32)["samples"]
        keep, m._h = m._h, h                       # run the same request on the C-loaded engine
        try:
            b = m.generate_debug(mel, True, 400, 100, seed=3, max_steps=64, precision=_native.PREC_F32)["samples"]
        finally:
            m._h = keep
        assert np.array_equal(a, b)
    finally:
        inference.unload()
        lib.wrnn_destroy(h)
