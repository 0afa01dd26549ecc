"""End-to-end throughput of the inference CLI (files in -> PNG files out) on synthetic JPEGs.

    python -m disentangledcolorization_b200.tools.cli_throughput [--images 256] [--batch 64] [--size 320x240]

Writes N smooth random colour JPEGs to a scratch directory, saves the synthetic checkpoint as a .pth.tar, runs
`inference.main` twice (the first run pays the CUDA context, cuDNN-free library load and workspace allocation) and prints
images/s of the second run: decode + resize + Lab on reader threads, H2D, forward, Lab->RGB on the device, D2H, PNG encode on
writer threads, all inside the timed region."""
import argparse
import json
import os
import shutil
import tempfile
import time

import numpy as np
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=256)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--size", default="320x240")
    ap.add_argument("--io_threads", type=int, default=min(16, os.cpu_count() or 4))
    args = ap.parse_args()
    import cv2
    from disentangledcolorization_b200 import inference, synth
    w, h = (int(v) for v in args.size.split("x"))
    tmp = tempfile.mkdtemp(prefix="disco_cli_")
    try:
        data = os.path.join(tmp, "data")
        os.makedirs(data)
        rng = np.random.default_rng(0)
        for i in range(args.images):
            small = rng.random((h // 16 + 2, w // 16 + 2, 3)).astype(np.float32)
            img = np.clip(cv2.resize(small, (w, h), interpolation=cv2.INTER_CUBIC) * 255, 0, 255).astype(np.uint8)
            cv2.imwrite(os.path.join(data, f"img{i:05d}.jpg"), img, [cv2.IMWRITE_JPEG_QUALITY, 92])
        ckpt = os.path.join(tmp, "model_last.pth.tar")
        torch.save({"state_dict": synth.make_state_dict(seed=0)}, ckpt)
        cwd = os.getcwd()
        os.chdir(tmp)
        res = {}
        try:
            for run in ("warm-up", "timed"):
                t0 = time.perf_counter()
                n, dt, dt_pipe = inference.main(["--data", data, "--checkpt", ckpt, "--name", run, "--batch", str(args.batch),
                                                 "--io_threads", str(args.io_threads)])
                res[run] = {"images": n, "seconds_load_to_done": dt, "seconds_pipeline": dt_pipe,
                            "seconds_total": time.perf_counter() - t0, "images_per_s": n / dt,
                            "images_per_s_pipeline": n / dt_pipe}
            n_png = len(os.listdir(os.path.join(tmp, "timed-anchor8")))
        finally:
            os.chdir(cwd)
        print(json.dumps({"cli_throughput": res, "png_written": n_png, "batch": args.batch, "io_threads": args.io_threads,
                          "source_size": args.size, "cores": os.cpu_count(),
                          "note": "seconds_load_to_done = checkpoint load + engine build + all images; seconds_pipeline = files in -> "
                                  "PNGs out with the model resident (includes the first forward's workspace allocation)"}))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
