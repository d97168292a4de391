"""Image-list part of the reference's argument helpers (args_util.py:7-27); the optimizer / phase parsers of that file
belong to the Keras training loop and are out of scope."""
from .data.voc_data_helpers import extract_img_data, get_img_names_from_set


def base_paths_to_imgs(base_path_str, img_set='trainval', do_flip=True):
    """Comma-separated VOC-layout roots -> list of shapes.Image, followed by their mirrored copies when `do_flip`."""
    imgs = []
    for path in base_path_str.split(','):
        imgs.extend(extract_img_data(path, name) for name in get_img_names_from_set(path, img_set))
    if do_flip:
        imgs += [img.horizontal_flip() for img in imgs]
    return imgs
