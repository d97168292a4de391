"""PASCAL-VOC style annotation loader (reference: data/voc_data_helpers.py:10-150), host side.

Same constants, same directory layout (`Annotations/ ImageSets/Main/ JPEGImages/`), same conventions: annotation
corners are 1-based and stored 0-based (`int(float(text)) - 1`), `difficult` is a bool, sizes come from the XML's
<size> node, and an image directory without an annotation file (KITTI test data) gets a minimal annotation built
from the image's own dimensions -- in memory here, the reference writes the XML next to the data.  Returns
`faster_rcnn_b200.shapes.Image` objects whose pixels are read lazily from `image_path`."""
import os
from xml.etree import ElementTree

from ..shapes import Box, GroundTruthBox, Image

ANNOTATIONS_DIR = 'Annotations'
IMAGESETS_DIR = os.path.join('ImageSets', 'Main')
IMAGES_DIR = 'JPEGImages'

VOC_CLASSES = ['aeroplane', 'bicycle', 'bird', 'boat', 'bottle', 'bus', 'car', 'cat', 'chair', 'cow', 'diningtable',
               'dog', 'horse', 'motorbike', 'person', 'pottedplant', 'sheep', 'sofa', 'train', 'tvmonitor']
VOC_CLASS_MAPPING = dict({c: i for i, c in enumerate(VOC_CLASSES)}, bg=20)
KITTI_CLASS_MAPPING = {'car': 0, 'person': 1, 'Cyclist': 2, 'DontCare': 3, 'Misc': 4, 'Person_sitting': 5, 'Tram': 6,
                       'Truck': 7, 'Van': 8, 'bg': 9}


def extract_img_data(base_path, img_num):
    """voc_data_helpers.py:68-129: one image's metadata (0-based GT corners, difficult flags) as a lazy Image."""
    annotations_path = os.path.join(base_path, ANNOTATIONS_DIR, img_num + '.xml')
    images_base = os.path.join(base_path, IMAGES_DIR)
    if not os.path.exists(annotations_path):
        import cv2
        image_path = os.path.join(images_base, img_num + '.png')
        raw = cv2.imread(image_path)
        if raw is None:
            raise IOError("no annotation and no image for %s under %s" % (img_num, base_path))
        return Image(img_num, raw.shape[1], raw.shape[0], [], image_path=image_path)
    annotation = ElementTree.parse(annotations_path).getroot()
    image_path = os.path.join(images_base, annotation.find('filename').text)
    size = annotation.find('size')
    width, height = int(size.find('width').text), int(size.find('height').text)
    gt_boxes = []
    for obj in annotation.findall('object'):
        bnd = obj.find('bndbox')
        xmin, xmax = int(float(bnd.find('xmin').text)) - 1, int(float(bnd.find('xmax').text)) - 1
        ymin, ymax = int(float(bnd.find('ymin').text)) - 1, int(float(bnd.find('ymax').text)) - 1
        difficult = int(obj.find('difficult').text) == 1
        gt_boxes.append(GroundTruthBox(obj.find('name').text, difficult, Box(xmin, ymin, xmax, ymax)))
    return Image(img_num, width, height, gt_boxes, image_path=image_path)


def get_img_names_from_set(base_path, set_name):
    """voc_data_helpers.py:132-138: image ids listed in ImageSets/Main/<set_name>.txt."""
    with open(os.path.join(base_path, IMAGESETS_DIR, set_name + '.txt')) as f:
        return [line.rstrip('\n') for line in f]
