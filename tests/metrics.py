"""Image-parity metrics of SURVEY.md §8(c): linear values, peak 1.0."""
import numpy as np


def compare(a: np.ndarray, b: np.ndarray, block: int = 16) -> dict:
    a = np.asarray(a, dtype=np.float64)[..., :3]
    b = np.asarray(b, dtype=np.float64)[..., :3]
    d = a - b
    mse = float((d * d).mean())
    h, w = a.shape[:2]
    hb, wb = h // block, w // block
    blk = lambda x: x[:hb * block, :wb * block].reshape(hb, block, wb, block, 3).mean(axis=(1, 3))
    return {
        "psnr": float(10 * np.log10(1.0 / max(mse, 1e-30))),
        "mae": [float(x) for x in np.abs(d).mean(axis=(0, 1))],
        "mean_diff": [float(x) for x in d.mean(axis=(0, 1))],
        "block_mae": float(np.abs(blk(a) - blk(b)).mean()) if hb and wb else 0.0,
        "block_max": float(np.abs(blk(a) - blk(b)).max()) if hb and wb else 0.0,
    }
