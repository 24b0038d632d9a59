"""`.pfe` reader/writer (harness plumbing around the hot path) and BASELINE config 1:
a 1024x1024 2-layer Normal-blend project flattened (CPU oracle here; the GPU run is in test_gpu_parity)."""
import struct

import numpy as np
import pytest

import fixtures as fx
from paintfe_b200 import pfe_io


def _config1_project():
    """L0 = test gradient, L1 = uniform-random RGBA; Normal; opacity 1.0 / 0.75."""
    rng = np.random.default_rng(0x5EED)
    l0 = fx.gradient(1024, 1024)
    l1 = rng.integers(0, 256, (1024, 1024, 4), dtype=np.uint8)
    l1[:200, :300, 3] = 0  # a transparent region: those chunks are not stored
    return pfe_io.PfeProject(1024, 1024, 1, [pfe_io.layer_from_flat("Background", l0),
                                            pfe_io.layer_from_flat("Layer 2", l1, opacity=0.75)]), (l0, l1)


def test_v1_roundtrip_and_layout():
    proj, (l0, l1) = _config1_project()
    raw = pfe_io.save_pfe_v1(proj)
    assert raw[8:12] == b"PFE1" and struct.unpack("<Q", raw[:8])[0] == 4  # bincode String: u64 length + bytes
    back = pfe_io.load_pfe_from_bytes(raw)
    assert (back.width, back.height, back.active_layer_index) == (1024, 1024, 1)
    assert [L.name for L in back.layers] == ["Background", "Layer 2"]
    assert back.layers[1].opacity == pytest.approx(0.75) and back.layers[1].blend_mode == 0
    assert np.array_equal(back.layers[0].to_flat(1024, 1024), l0)
    exp1 = l1.copy()
    exp1[:192, :256] = 0  # chunks wholly inside the transparent region are dropped (RGB lost), tiled_image.rs:82-97
    got1 = back.layers[1].to_flat(1024, 1024)
    assert np.array_equal(got1[192:], l1[192:]) and np.array_equal(got1[:192, :256], exp1[:192, :256])
    assert (0, 0) not in back.layers[1].chunks and back.layers[1].occupancy(1024, 1024)[0, 0] == 0
    assert pfe_io.save_pfe_v1(back) == raw  # byte-stable


def test_config1_flatten_via_oracle(oracle):
    proj, _ = _config1_project()
    back = pfe_io.load_pfe_from_bytes(pfe_io.save_pfe_v1(proj))
    flats = [L.to_flat(1024, 1024) for L in back.layers]
    out = oracle.flatten([oracle.make_layer(f, opacity=L.opacity, blend=L.blend_mode, visible=L.visible)
                          for f, L in zip(flats, back.layers)], 1024, 1024)
    assert out.shape == (1024, 1024, 4) and out[..., 3].min() == 255  # opaque background stays opaque
    assert np.array_equal(out[:100, :100], flats[0][:100, :100])      # transparent top region shows the gradient


def test_v0_and_v2_readers_and_validation():
    w, h = 70, 65
    rng = np.random.default_rng(1)
    flat = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    s = lambda t: struct.pack("<Q", len(t)) + t
    v0 = s(b"PFE0") + struct.pack("<IIQQ", w, h, 0, 1) + s(b"bg") + struct.pack("<BfB", 1, 0.5, 3) + s(flat.tobytes())
    p0 = pfe_io.load_pfe_from_bytes(v0)
    assert np.array_equal(p0.layers[0].to_flat(w, h), flat) and p0.layers[0].blend_mode == 3
    L = pfe_io.layer_from_flat("t", flat)
    body = b"".join(struct.pack("<IIQ", cx, cy, pfe_io.CHUNK_BYTES) + t.tobytes() for (cx, cy), t in sorted(L.chunks.items(), key=lambda k: (k[0][1], k[0][0])))
    v2 = (s(b"PFE2") + struct.pack("<IIQQ", w, h, 0, 1) + s(b"t") + struct.pack("<BfBB", 1, 1.0, 0, 1) +
          struct.pack("<Q", len(L.chunks)) + body + b"\x01" + s(b"textdata"))
    assert np.array_equal(pfe_io.load_pfe_from_bytes(v2).layers[0].to_flat(w, h), flat)
    for bad, msg in [(b"short", "too small"), (s(b"XXXX") + b"\0" * 8, "Unknown magic"), (s(b"PFE3") + b"\0" * 8, "zero"), (s(b"PFE3") + struct.pack("<IIQ", 5, 5, 0), "end of file"),
                     (s(b"PFE1") + struct.pack("<IIQQ", 0, 5, 0, 0), "zero"),
                     (s(b"PFE1") + struct.pack("<IIQQ", 30000, 5, 0, 0), "exceeds"),
                     (s(b"PFE1") + struct.pack("<IIQQ", 5, 5, 0, 300), "layers"),
                     (s(b"PFE1") + struct.pack("<IIQQ", 5, 5, 0, 1) + s(b"x") + struct.pack("<BfB", 1, 1.0, 0) +
                      struct.pack("<Q", 1) + struct.pack("<IIQ", 0, 0, 3) + b"abc", "expected")]:
        with pytest.raises(pfe_io.PfeError, match=msg):
            pfe_io.load_pfe_from_bytes(bad)


def test_v3_roundtrip_adjustment_layers_and_folders(oracle):
    """v3 (io.rs:171-208): folders, an adjustment layer and per-layer metadata survive a round trip, and
    the composite honours folder visibility and the adjustment layer (checked against the oracle)."""
    w, h = 130, 70
    rng = np.random.default_rng(2)
    bg = fx.random_rgba(rng, w, h, alpha="opaque")
    top = fx.random_rgba(rng, w, h)
    hidden = fx.random_rgba(rng, w, h)
    layers = [pfe_io.layer_from_flat("bg", bg), pfe_io.layer_from_flat("in hidden folder", hidden),
              pfe_io.PfeLayer("invert", True, 0.6, 0, adjustment=(3, ())),
              pfe_io.layer_from_flat("top", top, opacity=0.8, blend_mode=8),
              pfe_io.PfeLayer("mixer", True, 1.0, 0, adjustment=(4, tuple(float(v) for v in rng.uniform(-0.2, 1.1, 16)))),
              pfe_io.PfeLayer("exposure", True, 0.5, 0, adjustment=(1, (0.7,)))]
    layers[1].folder_id = 7
    layers[3].folder_id = 9
    proj = pfe_io.PfeProject(w, h, 3, layers, folders=[dict(id=7, name="off", visible=False, collapsed=True, insert_above_layer=None, color_index=2),
                                                      dict(id=9, name="on", visible=True, collapsed=False, insert_above_layer=1, color_index=None)],
                             next_layer_folder_id=10)
    raw = pfe_io.save_pfe_v3(proj)
    assert raw[8:12] == b"PFE3"
    back = pfe_io.load_pfe_from_bytes(raw)
    assert pfe_io.save_pfe_v3(back) == raw  # byte-stable, metadata tail carried verbatim
    assert [L.layer_type for L in back.layers] == [0, 0, 2, 0, 2, 2] and back.next_layer_folder_id == 10
    assert back.layers[2].adjustment == (3, ()) and back.layers[4].adjustment[0] == 4 and len(back.layers[4].adjustment[1]) == 16
    assert back.layers[5].adjustment[1][0] == pytest.approx(0.7)
    assert [back.layer_effectively_visible(i) for i in range(6)] == [True, False, True, True, True, True]
    o_layers = []
    active = np.zeros(((h + 63) // 64, (w + 63) // 64), np.uint8)
    for i, L in enumerate(back.layers):
        vis = back.layer_effectively_visible(i)
        if L.adjustment is not None:
            kind, prm = L.adjustment
            adj = (float(np.float32(2.0) ** np.float32(prm[0])),) if kind == 1 else prm
            o_layers.append(oracle.make_layer(None, opacity=L.opacity, visible=vis, kind=kind, adj=adj))
        else:
            o_layers.append(oracle.make_layer(L.to_flat(w, h), opacity=L.opacity, blend=L.blend_mode, visible=vis))
            if vis:
                active |= L.occupancy(w, h)
    out = oracle.flatten(o_layers, w, h, active=active)
    assert out.shape == (h, w, 4)
    # the hidden-folder layer must not show: same result with that layer removed altogether
    out2 = oracle.flatten([l for i, l in enumerate(o_layers) if i != 1], w, h, active=active)
    assert np.array_equal(out, out2)
