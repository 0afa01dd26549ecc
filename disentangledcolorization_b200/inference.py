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


def test_model(args):
    from . import basic, model as disco_model
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
    assert os.path.exists(args.checkpt), args.checkpt
    load_checkpoint(args.checkpt, net)
    print("-weight loaded successfully.")
    net = net.cuda().eval()
    start = time.time()
    n_done = 0
    pending = []

    def flush():
        nonlocal n_done
        if not pending:
            return
        grays = torch.cat([p[0] for p in pending]).cuda(non_blocking=True)
        abs_ = torch.cat([p[1] for p in pending]).cuda(non_blocking=True)
        sampled_T = 2 if args.diverse else 0
        out = net(grays, abs_, True, sampled_T)
        enhanced = out[2]
        if args.diverse:
            for no in range(3):
                lab = basic.tensor2array(torch.cat((grays, enhanced[no:no + 1]), dim=1))
                H, W = pending[0][3]
                lab = lab[:, :H, :W, :] if args.no_resize else lab
                save_lab_batch(lab, save_dir, [pending[0][2]], suffix="c%d" % no)
        else:
            lab = basic.tensor2array(torch.cat((grays, enhanced), dim=1))
            for i, (_, _, name, (H, W)) in enumerate(pending):
                one = lab[i:i + 1, :H, :W, :] if args.no_resize else lab[i:i + 1]
                save_lab_batch(one, save_dir, [name])
        n_done += len(pending)
        pending.clear()

    for img_path in img_list:
        fname = os.path.splitext(os.path.basename(img_path))[0] + ".png"
        print("-processing %s ..." % os.path.basename(img_path))
        gray, ab, hw = fetch_data(img_path, args.no_resize)
        if pending and (len(pending) >= args.batch or pending[0][0].shape != gray.shape or args.diverse):
            flush()
        pending.append((gray, ab, fname, hw))
        if len(pending) >= args.batch or args.diverse:
            flush()
    flush()
    print("-processed %d imgs. consumed %f sec" % (n_done, time.time() - start))


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
    return p


def main(argv=None):
    print("FLAG: %s" % datetime.datetime.now().strftime("%Y-%m-%d %H:%M:%S.%f"))
    args = build_parser().parse_args(argv)
    args.dense_pos = True                 # reference inference.py:165-166
    args.model = "AnchorColorProb"
    test_model(args)


if __name__ == "__main__":
    main()
