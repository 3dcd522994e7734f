"""Regenerates the committed fixtures under tests/golden/ (run in the build container, where
/root/reference is mounted; nothing at test time reads /root/reference).

  ref_book3_150.npy / ref_mixed_pdf_150.npy
        4x4 block means (150x150x3, float32, 8-bit sRGB scale) of the reference's own shipped
        renders final_images/book3.png (cornell_box at HEAD = config c5) and
        final_images/mixed_pdf.png (= config c2), and ref_cornell_smoke_150.npy of final_images/cornell_smoke.png
        (= config c3: deterministic geometry src/main.rs:514-601, scattering smoke = ISO 1/(4 pi), lights = empty).
        The oracle is pinned against these by PSNR.
  oracle_<cfg>[_lights].npz
        per-pixel mean and variance of the ORACLE (reference sampler mode, f64) at reduced
        resolution and 1024 spp: the "high-spp CPU render" the GPU images are accepted against.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
OUT = Path(__file__).resolve().parent

GOLDEN_RENDERS = {  # name: (config, variant, width, spp)
    "oracle_c1": ("c1", 0, 80, 1024),      # 5x5 blocks of the 400x225 config
    "oracle_c2": ("c2", 0, 120, 1024),
    "oracle_c3": ("c3", 0, 120, 1024),
    "oracle_c3_lights": ("c3", 1, 120, 1024),
    "oracle_c4": ("c4", 0, 160, 1024),     # 5x5 blocks of the 800x800 config
    "oracle_c4_lights": ("c4", 1, 160, 1024),
    "oracle_c5": ("c5", 0, 120, 1024),
}


def reference_pngs():
    from PIL import Image
    for png, out in (("book3.png", "ref_book3_150.npy"), ("mixed_pdf.png", "ref_mixed_pdf_150.npy"),
                     ("cornell_smoke.png", "ref_cornell_smoke_150.npy")):
        img = np.asarray(Image.open(f"/root/reference/final_images/{png}").convert("RGB")).astype(np.float64)
        assert img.shape == (600, 600, 3)
        blocks = img.reshape(150, 4, 150, 4, 3).mean(axis=(1, 3)).astype(np.float32)
        np.save(OUT / out, blocks)
        print(out, blocks.shape, blocks.mean())


def oracle_renders():
    from oracle import orc
    from surely_raytracing_b200.scenes import BuiltScene
    for name, (cfg, variant, width, spp) in GOLDEN_RENDERS.items():
        b = BuiltScene(cfg, width=width, spp=spp, variant=variant)
        o = orc.OracleScene(b)
        s, s2, st = o.render(sampler=orc.SAMPLER_REF, want_sumsq=True)
        n = o.info.spp_used
        mean = s / n
        var = np.maximum(s2 / n - mean ** 2, 0) * n / (n - 1)
        np.savez_compressed(OUT / f"{name}.npz", mean=mean.astype(np.float32), var=var.astype(np.float32),
                            spp=np.int32(n), width=np.int32(width), variant=np.int32(variant),
                            seg_per_path=np.float32(st["segments"] / st["paths"]))
        print(name, mean.shape, n, "mean", mean.mean(), "Mpaths/s", st["paths"] / st["wall_ms"] / 1e3)


if __name__ == "__main__":
    reference_pngs()
    oracle_renders()
