"""``imgaug`` is not installable offline; the reference imports it at module level (image_augmentation.py:7) and uses it only in
training augmentations.  Empty stand-in so that the TEST transform imports.  TEST INFRASTRUCTURE."""
