"""see ../__init__.py.  The reference CONSTRUCTS its training augmenters even for the test transform
(image_augmentation.py:145-147: MotionBlur, GaussianBlur, Sequential([OneOf(...)])) but never calls them with probability 0;
the stand-ins accept any arguments and refuse to be applied."""


class _Augmenter(object):
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        raise NotImplementedError("imgaug stand-in: training augmentations are out of scope")

    augment_image = augment_images = __call__


MotionBlur = GaussianBlur = OneOf = Sequential = Sometimes = JpegCompression = _Augmenter
