"""Image half of the loader, CPU side: (1) `GpuImageTransform.draw_params` consumes torch's generator exactly like the
reference transform (same crop boxes, same flips, same stream position afterwards) and (2) the resampling arithmetic the
GPU kernels implement - exercised here through its host restatement in the same library (`ecamp_image_resized_crop_host`,
test infrastructure: the product never calls it) - reproduces the bytes of the UNMODIFIED reference transform
(fixtures from oracle/make_image_golden.py) bit for bit, and Pillow itself on further random boxes."""
import ctypes
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from tests.image_frames import make_frame

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "image_pipeline.json")))


def host_resample(lib, frame, i, j, h, w, flip, out=448):
    crop = np.ascontiguousarray(frame[i:i + h, j:j + w])
    dst = np.empty((out, out), np.uint8)
    rc = lib.ecamp_image_resized_crop_host(crop.ctypes.data_as(ctypes.c_void_p), h, w, int(flip), out, dst.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return dst


def test_params_follow_the_reference_generator():
    from ecamp_b200.image_pipeline import GpuImageTransform
    t = GpuImageTransform(device="cpu")          # drawing parameters is host logic; no kernel runs here
    for c in GOLD["cases"]:
        torch.manual_seed(c["torch_seed"])
        i, j, h, w, flip = t.draw_params(c["H"], c["W"])
        assert (i, j, h, w, flip) == (c["i"], c["j"], c["h"], c["w"], c["flip"]), c
        assert float(torch.rand(1)) == c["next_rand"]      # the generator is where the reference leaves it


def test_resampling_matches_reference_fixtures(built_lib):
    lib = built_lib.lib()
    mean, std = torch.tensor(0.4721), torch.tensor(0.3037)
    for c in GOLD["cases"]:
        frame = make_frame(c["H"], c["W"], c["frame_seed"])
        out = host_resample(lib, frame, c["i"], c["j"], c["h"], c["w"], c["flip"])
        for y, x, v in c["samples"]:
            assert out[y, x] == v, (c["torch_seed"], y, x, int(out[y, x]), v)
        assert hashlib.sha256(out.tobytes()).hexdigest() == c["sha256_u8"], c["torch_seed"]
        f32 = torch.from_numpy(out).float().div(255).sub(mean).div(std)     # ToTensor + Normalize (pretrain_datasets.py:51-52)
        f32 = f32[None].expand(3, 448, 448).contiguous()
        assert hashlib.sha256(f32.numpy().tobytes()).hexdigest() == c["sha256_f32"], c["torch_seed"]


def test_resampling_matches_pillow_on_random_boxes(built_lib):
    PIL = pytest.importorskip("PIL.Image")
    lib = built_lib.lib()
    rs = np.random.RandomState(5)
    for n in range(12):
        H, W = int(rs.randint(60, 900)), int(rs.randint(60, 900))
        frame = make_frame(H, W, 20 + n)
        h, w = int(rs.randint(8, H + 1)), int(rs.randint(8, W + 1))
        i, j = int(rs.randint(0, H - h + 1)), int(rs.randint(0, W - w + 1))
        ref = PIL.fromarray(frame).convert("RGB").crop((j, i, j + w, i + h)).resize((448, 448), PIL.BICUBIC)
        ref = np.asarray(ref.convert("L"))
        out = host_resample(lib, frame, i, j, h, w, False)
        assert np.array_equal(out, ref), (H, W, i, j, h, w, int(np.abs(out.astype(int) - ref.astype(int)).max()))
