"""Drop-in for the reference's super-pixel CLI `main/spixelseg/inference.py` (same flags, same outputs):

    python -m disentangledcolorization_b200.spixel_inference --data DIR --checkpt .../checkpts/model_last.pth.tar --name result

For every image of --data (sorted; 256 x 256 like the reference, whose id grid is built for that size -- here any size that
is a multiple of 16): SpixelSeg forward -> winner-take-all super-pixel map (`split_spixels`, disco_spixel_ids) drawn on the
image, the colour image rebuilt from per-super-pixel mean colours (`poolfeat` -> `upfeat`, saved as `-recon`), and the gray
input (`-g`), written to <checkpoint dir>/../../<name>/ exactly as the reference names them (inference.py:39-108).
"""
import argparse
import datetime
import glob
import os
import time

import numpy as np
import torch


def fetch_data(img_path, need_padding=False):
    """reference main/spixelseg/inference.py:20-36 (its need_padding branch reads an undefined variable and is not built)."""
    import cv2
    if need_padding:
        raise NotImplementedError("fetch_data(need_padding=True) fails in the reference (NameError, inference.py:24); not built")
    bgr = cv2.imread(img_path, cv2.IMREAD_COLOR)
    if bgr is None:
        raise IOError(f"cannot read image {img_path}")
    bgr = np.array(bgr / 255.0, np.float32)
    lab = torch.from_numpy(cv2.cvtColor(bgr, cv2.COLOR_BGR2LAB).transpose((2, 0, 1)))
    bgr_t = torch.from_numpy(bgr.transpose((2, 0, 1)))
    gray = (lab[0:1] - 50.0) / 50.0
    color = lab[1:3] / 110.0
    return gray.unsqueeze(0), color.unsqueeze(0), (bgr_t * 2.0 - 1.0).unsqueeze(0)


def find_boundaries(label_img, background=0):
    """skimage.segmentation.find_boundaries(label_img, connectivity=1, mode='outer', background=0) as published
    (scikit-image 0.19-0.22), restated with scipy.ndimage; scikit-image is not part of this image, so this restatement is not
    pinned against it (the reference calls it through mark_boundaries, utils/util.py:116)."""
    from scipy import ndimage as ndi
    label_img = np.asarray(label_img)
    cross = ndi.generate_binary_structure(2, 1)
    boundaries = ndi.grey_dilation(label_img, footprint=cross) != ndi.grey_erosion(label_img, footprint=cross)
    max_label = np.iinfo(label_img.dtype).max
    background_image = label_img == background
    full = ndi.generate_binary_structure(2, 2)
    inverted_background = np.array(label_img, copy=True)
    inverted_background[background_image] = max_label
    adjacent_objects = (ndi.grey_dilation(label_img, footprint=full) != ndi.grey_erosion(inverted_background, footprint=full)) & ~background_image
    return boundaries & (background_image | adjacent_objects)


def mark_boundaries(image, label_img, color=(1, 1, 0)):
    """skimage.segmentation.mark_boundaries(image, label_img, color, mode='outer') (same caveat as find_boundaries)."""
    marked = np.array(image, dtype=np.float64, copy=True)
    marked[find_boundaries(label_img)] = color
    return marked


def save_markedSP_from_batch(img_batch, spix_batch, save_dir, filename_list, batch_no=-1, suffix=None):
    """reference utils/util.py:109-122."""
    from PIL import Image
    N = img_batch.shape[0]
    for i in range(N):
        norm_image = img_batch[i] * 0.5 + 0.5
        marked = mark_boundaries(norm_image, spix_batch[i, :, :, 0].astype(int), color=(1, 1, 1))
        save_name = filename_list[i] if batch_no == -1 else "%05d.png" % (batch_no * N + i)
        save_name = save_name.replace(".png", "-%s.png" % suffix) if suffix else save_name
        Image.fromarray((marked * 255.0).astype(np.uint8)).save(os.path.join(save_dir, save_name), "PNG")


def save_images_from_batch(img_batch, save_dir, filename_list, batch_no=-1, suffix=None):
    """reference utils/util.py:56-76 ([-1,1] -> [0,255]; 3-channel or single-channel)."""
    from PIL import Image
    N, H, W, Cc = img_batch.shape
    for i in range(N):
        arr = 127.5 * (img_batch[i] + 1.0) if Cc == 3 else 127.5 * (img_batch[i, :, :, 0] + 1.0)
        save_name = filename_list[i] if batch_no == -1 else "%05d.png" % (batch_no * N + i)
        save_name = save_name.replace(".png", "-%s.png" % suffix) if suffix else save_name
        Image.fromarray(arr.astype(np.uint8)).save(os.path.join(save_dir, save_name), "PNG")


def test_model(data_dir, model_name, sp_size, checkpt_path, name, precision="bf16"):
    """reference main/spixelseg/inference.py:39-108."""
    from . import basic, model as disco_model
    from .inference import load_checkpoint, save_lab_batch
    print("@Inference: [%s] (spixel-size=%d)" % (model_name, sp_size))
    root_dir = os.path.abspath(os.path.join(checkpt_path, "..", ".."))
    save_dir = os.path.join(root_dir, name)
    os.makedirs(save_dir, exist_ok=True)
    print("-loading dir:%s" % data_dir)
    print("-saving dir:%s" % save_dir)
    if model_name != "SpixelSeg":
        raise ValueError("only --model SpixelSeg exists in the reference's model.py")
    spix_model = disco_model.SpixelSeg(inChannel=1, outChannel=9, batchNorm=True)
    assert os.path.exists(checkpt_path), checkpt_path
    load_checkpoint(checkpt_path, spix_model)
    spix_model.net.precision = precision
    spix_model = spix_model.cuda().eval()
    print("-weight loaded successfully.")
    img_list = sorted(glob.glob(os.path.join(data_dir, "*.*")))
    nn_, start = 0, time.time()
    for img_pth in img_list:
        file_name = os.path.split(img_pth)[1]
        print("-processing %s ..." % file_name)
        input_gray, input_color, input_bgr = fetch_data(img_pth, False)
        input_gray, input_color, input_bgr = input_gray.cuda(), input_color.cuda(), input_bgr.cuda()
        with torch.no_grad():
            pred_probs = spix_model(input_gray)
            pred_spixel_map = basic.split_spixels(pred_probs, sp_size)                      # inference.py:67-75,96
            spix_color = basic.poolfeat(input_color, pred_probs, sp_size, sp_size)           # :104
            recon_color = basic.upfeat(spix_color, pred_probs, sp_size, sp_size)             # :105
        save_markedSP_from_batch(basic.tensor2array(input_bgr), basic.tensor2array(pred_spixel_map.float()), save_dir, [file_name], -1)
        save_lab_batch(basic.tensor2array(torch.cat((input_gray, recon_color), dim=1)), save_dir, [file_name], suffix="recon")
        save_images_from_batch(basic.tensor2array(input_gray), save_dir, [file_name], -1, suffix="g")
        nn_ += 1
    print("-processed %d imgs. consumed %f sec" % (nn_, time.time() - start))
    return nn_


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument("--name", type=str, default="result", help="save dir name")
    p.add_argument("--data", type=str, default="../../../0DataZoo/Dataset_C/VOC2012/Val/target", help="path of images")
    p.add_argument("--psize", default="16", type=int, help="super-pixel size")
    p.add_argument("--model", type=str, default="SpixelSeg", help="which model to use")
    p.add_argument("--checkpt", type=str, default="../../Saved/spixG2C_16/checkpts/model_last.pth.tar", help="path of weight")
    p.add_argument("--precision", default="bf16", choices=["bf16", "fp32"], help="[extension] bf16 tensor-core or fp32 exact path")
    return p


def main(argv=None):
    print("FLAG: %s" % datetime.datetime.now().strftime("%Y-%m-%d %H:%M:%S.%f"))
    args = build_parser().parse_args(argv)
    return test_model(args.data, args.model, args.psize, args.checkpt, args.name, args.precision)


if __name__ == "__main__":
    main()
