"""Drop-in for the reference CLI `main/colorizer/inference.py` (same flags, same I/O contract):

    python -m disentangledcolorization_b200.inference --data DIR --checkpt model_last.pth.tar --name out \
           [--n_clusters 8] [--no_resize] [--seed 130] [--psize 16] [--diverse] [--random_hint] ...

Reads every image of --data (sorted), converts to Lab with OpenCV exactly as `fetch_data`
(inference.py:23-42: resize to 256x256 INTER_LINEAR, or edge-pad to a multiple of 16 with --no_resize),
runs the B200 forward with `test_mode=True`, and writes <name>-anchor<K>/<file>.png (Lab -> RGB through OpenCV,
like util.save_normLabs_from_batch, utils/util.py:91-106).  Extensions: --batch N groups images of equal size
into one forward (the reference processes one image per call), --precision {bf16,fp32}.
"""
import argparse
import collections
import datetime
import glob
import os
import time

import numpy as np
import torch


def fetch_data(img_path, org_size=True):
    """reference main/colorizer/inference.py:23-42."""
    import cv2
    bgr = cv2.imread(img_path, cv2.IMREAD_COLOR)
    if bgr is None:
        raise IOError(f"cannot read image {img_path}")
    rgb = cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB)
    H, W = rgb.shape[:2]
    if org_size:
        if H % 16 != 0 or W % 16 != 0:       # note: pads a full extra 16 on an already aligned side, like the reference
            rgb = np.pad(rgb, ((0, 16 - H % 16), (0, 16 - W % 16), (0, 0)), mode="edge")
    else:
        rgb = cv2.resize(rgb, (256, 256), interpolation=cv2.INTER_LINEAR)
    rgb = np.array(rgb / 255.0, np.float32)
    lab = cv2.cvtColor(rgb, cv2.COLOR_RGB2LAB)
    lab_t = torch.from_numpy(lab.transpose((2, 0, 1)))
    gray = (lab_t[0:1] - 50.0) / 50.0
    ab = lab_t[1:3] / 110.0
    return gray.unsqueeze(0), ab.unsqueeze(0), (H, W)


def save_lab_batch(lab_nhwc, save_dir, names, suffix=None):
    """normalised Lab (N,H,W,3) -> RGB PNGs (reference utils/util.py:91-106)."""
    import cv2
    from PIL import Image
    lab = lab_nhwc.copy()
    lab[..., 0] = lab[..., 0] * 50.0 + 50.0
    lab[..., 1:3] = lab[..., 1:3] * 110.0
    for i in range(lab.shape[0]):
        rgb = cv2.cvtColor(lab[i], cv2.COLOR_LAB2RGB)
        name = names[i].replace(".png", f"-{suffix}.png") if suffix else names[i]
        Image.fromarray((rgb * 255.0).astype(np.uint8)).save(os.path.join(save_dir, name), "PNG")


def load_checkpoint(checkpt_path, model):
    """reference main/utils_train.py:140-156 (inference use)."""
    data = torch.load(checkpt_path, map_location=torch.device("cpu"))
    model.load_state_dict(data["state_dict"])
    return model


def _fetch_np(img_path, no_resize):
    """fetch_data for the reader threads: numpy (1,H,W) gray, (2,H,W) ab and the original size (cv2 releases the GIL)."""
    gray, ab, hw = fetch_data(img_path, no_resize)
    return gray[0].numpy(), ab[0].numpy(), hw


_U8_TO_UNIT = None


def _fetch_into(img_path, gray_out, ab_out):
    """`fetch_data(img_path, org_size=False)` written straight into (1,256,256) / (2,256,256) float32 views of a pinned
    batch buffer, bit-identical to it (tests/test_cli_cpu.py) but with every heavy step inside OpenCV / numpy loops that
    release the GIL: the uint8 -> [0,1] step `np.array(rgb / 255.0, np.float32)` becomes a 256-entry table lookup of the
    same float64-divided, float32-rounded values."""
    import cv2
    global _U8_TO_UNIT
    if _U8_TO_UNIT is None:
        _U8_TO_UNIT = np.array(np.arange(256) / 255.0, np.float32)
    bgr = cv2.imread(img_path, cv2.IMREAD_COLOR)
    if bgr is None:
        raise IOError(f"cannot read image {img_path}")
    H, W = bgr.shape[:2]
    rgb = cv2.resize(cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB), (256, 256), interpolation=cv2.INTER_LINEAR)
    lab = cv2.cvtColor(cv2.LUT(rgb, _U8_TO_UNIT), cv2.COLOR_RGB2LAB)
    L, A, B = cv2.split(lab)                                  # contiguous planes: the arithmetic below runs at memcpy speed
    np.subtract(L, np.float32(50.0), out=gray_out[0])
    np.divide(gray_out[0], np.float32(50.0), out=gray_out[0])
    np.divide(A, np.float32(110.0), out=ab_out[0])
    np.divide(B, np.float32(110.0), out=ab_out[1])
    return (H, W)


def _save_rgb(path, rgb):
    """PNG writer for the writer threads: the native encoder of the C ABI (disco_host_png_write, csrc/png_host.cu; the call
    releases the GIL).  The decoded pixels are what the reference's PIL save stores (tests/test_cli_cpu.py); PNG encoding is
    what bounded the CLI once the forward ran on the GPU (PIL level 6 = 18-26 ms, OpenCV level 1 =
    7.7 ms, this writer 1.3 ms per 256 x 256 image on one core)."""
    from . import _lib
    _lib.png_write(path, rgb)


def test_model(args):
    """reference main/colorizer/inference.py:56-139, restructured as a pipeline: reader threads decode /
    resize / convert to Lab, images of equal size are grouped (in sorted order, so the host RNG is consumed per image
    exactly as one image per forward would consume it) into batches of --batch, `ColorizePipeline` overlaps H2D, forward and
    D2H of neighbouring batches, Lab -> RGB uint8 runs on the device (disco_lab2rgb_u8), writer threads encode the PNGs."""
    import concurrent.futures as cf
    import ctypes as C
    from . import _lib, model as disco_model
    from .pipeline import ColorizePipeline
    print("@Inference: [%s] (spixel-size=%d)" % (args.model, args.psize))
    np.random.seed(args.seed)
    torch.manual_seed(args.seed)
    torch.cuda.manual_seed(args.seed)
    save_dir = os.path.abspath(args.name + "-anchor%d" % args.n_clusters)
    os.makedirs(save_dir, exist_ok=True)
    img_list = sorted(glob.glob(os.path.join(args.data, "*.*")))
    print("-data dir (%d images):%s" % (len(img_list), args.data))
    print("-saving dir:%s" % save_dir)
    net = disco_model.AnchorColorProb(inChannel=1, outChannel=313, sp_size=args.psize, d_model=args.d_model,
                                      use_dense_pos=args.dense_pos, spix_pos=args.spix_pos,
                                      learning_pos=args.learning_pos, n_clusters=args.n_clusters,
                                      random_hint=args.random_hint, hint2regress=args.hint2regress, enhanced=True,
                                      precision=args.precision)
    start = time.time()
    assert os.path.exists(args.checkpt), args.checkpt
    load_checkpoint(args.checkpt, net)
    print("-weight loaded successfully.")
    net = net.cuda().eval()
    net.batched_diverse = True
    t_ready = time.time()
    dev = torch.device("cuda", torch.cuda.current_device())
    handle = _lib.Handle.get(dev.index)
    sampled_T = 2 if args.diverse else 0
    n_var = 3 if args.diverse else 1
    n_threads = max(1, args.io_threads)
    import cv2
    cv_threads = cv2.getNumThreads()
    cv2.setNumThreads(1)                                  # parallelism comes from the reader / writer pools, not from inside OpenCV
    readers = cf.ThreadPoolExecutor(n_threads)
    writers = cf.ThreadPoolExecutor(n_threads)
    window = max(2 * args.batch, 2 * n_threads)          # decoded images in flight
    metas, writes = [], []                                # per batch: [(file name, (H, W)), ...]; pending PNG writes

    def batches_resized():
        """Default mode: every image becomes 256 x 256, so image k's place (batch k // B, row k % B) is known before it is
        decoded and the reader threads fill the pinned batch buffers themselves; the main thread only waits for a batch's
        futures.  A ring of 5 buffer sets: 2 in the GPU pipeline, 2 being filled, 1 spare; a set is refilled only after the
        H2D copy of its previous batch has completed (event recorded by the pipeline)."""
        B, n, R, look = args.batch, len(img_list), 5, 2
        nb = (n + B - 1) // B
        sets = [dict(gray=torch.empty(B, 1, 256, 256).pin_memory(), ab=torch.empty(B, 2, 256, 256).pin_memory(), ev=None)
                for _ in range(min(R, max(nb, 1)))]

        def submit(bi):
            st = sets[bi % len(sets)]
            if st["ev"] is not None:
                st["ev"].synchronize()
            g, a = st["gray"].numpy(), st["ab"].numpy()
            futs = []
            for k in range(bi * B, min(n, (bi + 1) * B)):
                print("-processing %s ..." % os.path.basename(img_list[k]))
                futs.append(readers.submit(_fetch_into, img_list[k], g[k - bi * B], a[k - bi * B]))
            return futs

        pend = {bi: submit(bi) for bi in range(min(look, nb))}
        for bi in range(nb):
            if bi + look < nb:
                pend[bi + look] = submit(bi + look)
            hws = [f.result() for f in pend.pop(bi)]
            st = sets[bi % len(sets)]
            st["ev"] = torch.cuda.Event()
            names = [os.path.splitext(os.path.basename(img_list[bi * B + i]))[0] + ".png" for i in range(len(hws))]
            metas.append(list(zip(names, hws)))
            yield st["gray"][:len(hws)], st["ab"][:len(hws)], st["ev"]

    def batches():
        """--no_resize: consecutive images of equal (padded) size, at most --batch per forward, staged in pinned memory."""
        pending = collections.deque()
        it = iter(img_list)

        def refill():
            while len(pending) < window:
                path = next(it, None)
                if path is None:
                    return
                print("-processing %s ..." % os.path.basename(path))
                pending.append((path, readers.submit(_fetch_np, path, args.no_resize)))

        refill()
        group = []
        while pending or group:
            item = None
            if pending:
                path, fut = pending.popleft()
                refill()
                g, a, hw = fut.result()
                item = (os.path.splitext(os.path.basename(path))[0] + ".png", g, a, hw)
            if group and (item is None or len(group) >= args.batch or group[0][1].shape != item[1].shape):
                gray = torch.empty(len(group), 1, *group[0][1].shape[1:]).pin_memory()
                ab = torch.empty(len(group), 2, *group[0][1].shape[1:]).pin_memory()
                for i, (_, g_, a_, _) in enumerate(group):
                    gray[i] = torch.from_numpy(g_)
                    ab[i] = torch.from_numpy(a_)
                metas.append([(nm, hw_) for nm, _, _, hw_ in group])
                yield gray, ab
                group = []
            if item is not None:
                group.append(item)

    def to_rgb(out, gray, ab):
        """compute stream, right after the forward: Lab -> RGB uint8 for every variant of every image of the batch"""
        pred = out[2]
        B, _, H, W = gray.shape
        g = gray if n_var == 1 else gray.repeat(n_var, 1, 1, 1)
        rgb = torch.empty(pred.shape[0], H, W, 3, dtype=torch.uint8, device=pred.device)
        _lib.check(handle.lib.disco_lab2rgb_u8(handle.h, C.c_void_p(g.data_ptr()), C.c_void_p(pred.data_ptr()), pred.shape[0], H, W,
                                               H, W, C.c_void_p(rgb.data_ptr()), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
                   "disco_lab2rgb_u8")
        return rgb

    def _save_many(items):
        for path, img in items:
            _save_rgb(path, img)

    def on_result(i, host):
        """host thread, batch i has landed in pinned memory: take a private copy (the slot buffer is reused two steps
        later), crop the padding off, hand the images to the writer threads in chunks"""
        arr = host.numpy().copy()
        B = len(metas[i])
        items = []
        for v in range(n_var):
            for k, (name, (H, W)) in enumerate(metas[i]):
                img = arr[v * B + k]
                img = img[:H, :W] if args.no_resize else img     # a view: the encoder takes a row stride
                out_name = name.replace(".png", "-c%d.png" % v) if args.diverse else name
                items.append((os.path.join(save_dir, out_name), img))
        chunk = max(1, len(items) // (2 * n_threads))
        for c in range(0, len(items), chunk):
            writes.append(writers.submit(_save_many, items[c:c + chunk]))

    pipe = ColorizePipeline(net, device=dev, sampled_T=sampled_T)
    pipe.run(batches() if args.no_resize else batches_resized(), post=to_rgb, on_result=on_result, keep="none")
    for w in writes:
        w.result()
    readers.shutdown()
    writers.shutdown()
    cv2.setNumThreads(cv_threads)
    n_done = sum(len(m) for m in metas)
    dt = time.time() - start
    print("-processed %d imgs. consumed %f sec" % (n_done, dt))
    return n_done, dt, time.time() - t_ready


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument("--name", type=str, default="test", help="save dir name")
    p.add_argument("--seed", default=130, type=int, help="random seed")
    p.add_argument("--psize", default=16, type=int, help="super-pixel size")
    p.add_argument("--data", type=str, default="../../../0DataZoo/Dataset_C/VOC2012/Val/target", help="path of images")
    p.add_argument("--model", type=str, default="AnchorColorProb", help="which model to use")
    p.add_argument("--checkpt", type=str, default="../../Saved/colorProb/checkpts/model_last.pth.tar", help="path of weight")
    p.add_argument("--n_enc", default=3, type=int, help="number of encoder layers")
    p.add_argument("--n_dec", default=6, type=int, help="number of decoder layers")
    p.add_argument("--d_model", default=64, type=int, help="feature dimension of transformer")
    p.add_argument("--dense_pos", action="store_true", default=False, help="use pos encoding at each SA block")
    p.add_argument("--spix_pos", action="store_true", default=False, help="use pos of spixel centroid")
    p.add_argument("--learning_pos", action="store_true", default=False, help="learnable pos embedding")
    p.add_argument("--hint2regress", action="store_true", default=False, help="predict ab values from hint")
    p.add_argument("--n_clusters", default=8, type=int, help="number of color clusters")
    p.add_argument("--random_hint", action="store_true", default=False, help="sample anchors randomly")
    p.add_argument("--no_resize", action="store_true", default=False, help="input the original resolution")
    p.add_argument("--diverse", action="store_true", default=False, help="use pixel-level enhancement or not")
    # extensions
    p.add_argument("--batch", default=1, type=int, help="[extension] images per forward (equal sizes only)")
    p.add_argument("--precision", default="bf16", choices=["bf16", "fp32"], help="[extension] bf16 tensor-core or fp32 exact path")
    p.add_argument("--io_threads", default=min(16, os.cpu_count() or 4), type=int,
                   help="[extension] reader (decode/resize/Lab) and writer (PNG) threads around the GPU pipeline")
    return p


def main(argv=None):
    print("FLAG: %s" % datetime.datetime.now().strftime("%Y-%m-%d %H:%M:%S.%f"))
    args = build_parser().parse_args(argv)
    args.dense_pos = True                 # reference inference.py:165-166
    args.model = "AnchorColorProb"
    return test_model(args)


if __name__ == "__main__":
    main()
