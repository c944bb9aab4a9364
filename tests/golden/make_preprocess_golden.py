"""Generates tests/golden/preprocess.npz: golden vectors for the test-time frame transform, produced by the third-party
code the reference calls (torchvision F.resize on a PIL image = Pillow's ImagingResample, ToTensor, F.normalize;
siammot/data/adapters/augmentation/image_augmentation.py:44-46, build_augmentation.py:52-66).

Run in the authoring container:  python tests/golden/make_preprocess_golden.py
Pillow version used is stored in the file (the reference pins 10.0.1; the resampling routine is unchanged since)."""
import os

import numpy as np
import PIL
import torch
from PIL import Image
from torchvision.transforms import functional as TF

HERE = os.path.dirname(os.path.abspath(__file__))
# (seed, frame h, w, output oh, ow, mean, std, to_bgr255)
CASES = [
    (1, 90, 160, 64, 96, (0.485, 0.456, 0.406), (0.229, 0.224, 0.225), False),      # both passes, down
    (2, 45, 80, 96, 160, (0.485, 0.456, 0.406), (0.229, 0.224, 0.225), False),      # both passes, up
    (3, 72, 128, 64, 128, (102.9801, 115.9465, 122.7717), (1.0, 1.0, 1.0), True),   # vertical pass only (720p -> 704 rule), BGR255
    (4, 64, 100, 64, 96, (0.5, 0.5, 0.5), (0.25, 0.25, 0.25), False),               # horizontal pass only
]


def frame(seed, h, w):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    img[: h // 3] = (img[: h // 3] // 128) * 255
    return img


def main():
    out = {"pillow_version": np.array(PIL.__version__)}
    for (seed, h, w, oh, ow, mean, std, bgr) in CASES:
        img = frame(seed, h, w)
        resized = TF.resize(Image.fromarray(img, "RGB"), (oh, ow))
        t = TF.to_tensor(resized)
        if bgr:
            t = t[[2, 1, 0]] * 255
        t = TF.normalize(t, mean=list(mean), std=list(std))
        out["resized_%d" % seed] = np.asarray(resized)
        out["tensor_%d" % seed] = t.numpy()
    np.savez_compressed(os.path.join(HERE, "preprocess.npz"), **out)
    print("wrote", os.path.join(HERE, "preprocess.npz"))


if __name__ == "__main__":
    main()
