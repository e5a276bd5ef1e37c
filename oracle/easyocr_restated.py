"""CPU oracle (TEST INFRASTRUCTURE ONLY) for EasyOCR 1.7.2's `Reader.readtext` as BB-OCR calls it.

PARITY UNPINNED.  The arithmetic of this stage lives in the third-party package easyocr==1.7.2
(/root/reference/pipeline_demo/requirements.txt:7), which is neither vendored under /root/reference nor installed
in this image, and whose checkpoints (craft_mlt_25k.pth, english_g2.pth) cannot be downloaded (no network).  This
file restates the published algorithm of that release (easyocr/easyocr.py, detection.py, craft.py, craft_utils.py,
imgproc.py, utils.py, recognition.py, model/modules.py, model/vgg_model.py, config.py) function by function, anchored
on the reference's call sites:
    pipeline_demo/extractor/enhanced_extractor.py:153   easyocr.Reader(["en"], gpu=use_gpu)
    pipeline_demo/extractor/enhanced_extractor.py:520   reader.readtext(path, paragraph=False, batch_size=1, workers=0)
    pipeline_components/img_to_json/ocr_testing/ocr_engines/test_easyocr.py:20-23,50-53  (bbox, text, prob) consumer
Structural evidence: parameter counts 20 770 466 (CRAFT) and 3 781 345 (CRNN) equal the published
checkpoints'; state-dict key names follow the upstream modules so real .pth files load unchanged.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
The FP32 oracle corresponds to `Reader(['en'], gpu=False, quantize=False)`; EasyOCR's default CPU path additionally
int8-quantises LSTM/Linear (quantize=True) -- available here as `Reader(quantize=True)` for the timing baseline only.
"""
from __future__ import annotations

import math

import cv2
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from PIL import Image

# easyocr/config.py : recognition_models['gen2']['english_g2']
SYMBOLS = "0123456789!\"#$%&'()*+,-./:;<=>?@[\\]^_`{|}~ €"
CHARACTERS = SYMBOLS + "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz"
IMG_H = 64


# ======================================================================================================================
# Networks (easyocr/craft.py, easyocr/model/modules.py, easyocr/model/vgg_model.py)
# ======================================================================================================================

def _vgg16_bn_features():
    """torchvision.models.vgg16_bn().features, re-stated (cfg 'D', batch_norm=True, ReLU(inplace=True))."""
    cfg = [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512, "M"]
    layers, c = [], 3
    for v in cfg:
        if v == "M":
            layers.append(nn.MaxPool2d(2, 2))
        else:
            layers += [nn.Conv2d(c, v, 3, padding=1), nn.BatchNorm2d(v), nn.ReLU(inplace=True)]
            c = v
    return nn.Sequential(*layers)


class VGG16BN(nn.Module):
    """easyocr/model/modules.py::vgg16_bn -- slices 0:12, 12:19, 19:29, 29:39 + dilated fc6/fc7.

    NB the first module of slice2/3/4 is an *in-place* ReLU, so the tensors aliased as relu2_2 / relu3_2 / relu4_3
    are rectified by the time the decoder reads them; relu5_3 (followed by a MaxPool) stays the raw BN output.
    """

    def __init__(self):
        super().__init__()
        feats = _vgg16_bn_features()
        self.slice1, self.slice2, self.slice3 = nn.Sequential(), nn.Sequential(), nn.Sequential()
        self.slice4, self.slice5 = nn.Sequential(), nn.Sequential()
        for x in range(12):
            self.slice1.add_module(str(x), feats[x])
        for x in range(12, 19):
            self.slice2.add_module(str(x), feats[x])
        for x in range(19, 29):
            self.slice3.add_module(str(x), feats[x])
        for x in range(29, 39):
            self.slice4.add_module(str(x), feats[x])
        self.slice5 = nn.Sequential(nn.MaxPool2d(3, 1, 1), nn.Conv2d(512, 1024, 3, padding=6, dilation=6),
                                    nn.Conv2d(1024, 1024, 1))

    def forward(self, X):
        h = self.slice1(X); r22 = h
        h = self.slice2(h); r32 = h
        h = self.slice3(h); r43 = h
        h = self.slice4(h); r53 = h
        h = self.slice5(h)
        return h, r53, r43, r32, r22


class _DoubleConv(nn.Module):
    def __init__(self, in_ch, mid_ch, out_ch):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(in_ch + mid_ch, mid_ch, 1), nn.BatchNorm2d(mid_ch), nn.ReLU(inplace=True),
                                  nn.Conv2d(mid_ch, out_ch, 3, padding=1), nn.BatchNorm2d(out_ch), nn.ReLU(inplace=True))

    def forward(self, x):
        return self.conv(x)


class CRAFT(nn.Module):
    """easyocr/craft.py::CRAFT.  Output: (N, H/2, W/2, 2) score maps [text, link], and the 32-ch feature."""

    def __init__(self):
        super().__init__()
        self.basenet = VGG16BN()
        self.upconv1 = _DoubleConv(1024, 512, 256)
        self.upconv2 = _DoubleConv(512, 256, 128)
        self.upconv3 = _DoubleConv(256, 128, 64)
        self.upconv4 = _DoubleConv(128, 64, 32)
        self.conv_cls = nn.Sequential(
            nn.Conv2d(32, 32, 3, padding=1), nn.ReLU(inplace=True), nn.Conv2d(32, 32, 3, padding=1), nn.ReLU(inplace=True),
            nn.Conv2d(32, 16, 3, padding=1), nn.ReLU(inplace=True), nn.Conv2d(16, 16, 1), nn.ReLU(inplace=True),
            nn.Conv2d(16, 2, 1))

    def forward(self, x):
        s = self.basenet(x)
        y = self.upconv1(torch.cat([s[0], s[1]], 1))
        y = F.interpolate(y, size=s[2].shape[2:], mode="bilinear", align_corners=False)
        y = self.upconv2(torch.cat([y, s[2]], 1))
        y = F.interpolate(y, size=s[3].shape[2:], mode="bilinear", align_corners=False)
        y = self.upconv3(torch.cat([y, s[3]], 1))
        y = F.interpolate(y, size=s[4].shape[2:], mode="bilinear", align_corners=False)
        feature = self.upconv4(torch.cat([y, s[4]], 1))
        y = self.conv_cls(feature)
        return y.permute(0, 2, 3, 1), feature


class _VGGFeatureExtractor(nn.Module):
    def __init__(self, input_channel=1, output_channel=256):
        super().__init__()
        oc = [output_channel // 8, output_channel // 4, output_channel // 2, output_channel]
        self.ConvNet = nn.Sequential(
            nn.Conv2d(input_channel, oc[0], 3, 1, 1), nn.ReLU(True), nn.MaxPool2d(2, 2),
            nn.Conv2d(oc[0], oc[1], 3, 1, 1), nn.ReLU(True), nn.MaxPool2d(2, 2),
            nn.Conv2d(oc[1], oc[2], 3, 1, 1), nn.ReLU(True),
            nn.Conv2d(oc[2], oc[2], 3, 1, 1), nn.ReLU(True), nn.MaxPool2d((2, 1), (2, 1)),
            nn.Conv2d(oc[2], oc[3], 3, 1, 1, bias=False), nn.BatchNorm2d(oc[3]), nn.ReLU(True),
            nn.Conv2d(oc[3], oc[3], 3, 1, 1, bias=False), nn.BatchNorm2d(oc[3]), nn.ReLU(True), nn.MaxPool2d((2, 1), (2, 1)),
            nn.Conv2d(oc[3], oc[3], 2, 1, 0), nn.ReLU(True))

    def forward(self, x):
        return self.ConvNet(x)


class _BidirectionalLSTM(nn.Module):
    def __init__(self, input_size, hidden_size, output_size):
        super().__init__()
        self.rnn = nn.LSTM(input_size, hidden_size, bidirectional=True, batch_first=True)
        self.linear = nn.Linear(hidden_size * 2, output_size)

    def forward(self, x):
        r, _ = self.rnn(x)
        return self.linear(r)


class CRNN(nn.Module):
    """easyocr/model/vgg_model.py::Model(input_channel=1, output_channel=256, hidden_size=256, num_class=97)."""

    def __init__(self, input_channel=1, output_channel=256, hidden_size=256, num_class=len(CHARACTERS) + 1):
        super().__init__()
        self.FeatureExtraction = _VGGFeatureExtractor(input_channel, output_channel)
        self.AdaptiveAvgPool = nn.AdaptiveAvgPool2d((None, 1))
        self.SequenceModeling = nn.Sequential(_BidirectionalLSTM(output_channel, hidden_size, hidden_size),
                                              _BidirectionalLSTM(hidden_size, hidden_size, hidden_size))
        self.Prediction = nn.Linear(hidden_size, num_class)

    def forward(self, x, text=None):
        v = self.FeatureExtraction(x)
        v = self.AdaptiveAvgPool(v.permute(0, 3, 1, 2)).squeeze(3)
        c = self.SequenceModeling(v)
        return self.Prediction(c.contiguous())


# ======================================================================================================================
# Detection (easyocr/imgproc.py, detection.py, craft_utils.py)
# ======================================================================================================================

def reformat_input(image):
    """easyocr/utils.py::reformat_input for the input kinds BB-OCR and its legacy callers use."""
    if isinstance(image, str):
        img_cv_grey = cv2.imread(image, cv2.IMREAD_GRAYSCALE)
        bgr = cv2.imread(image, cv2.IMREAD_COLOR)        # upstream: skimage.io.imread (RGB); cv2 decode + swap here
        if bgr is None:
            raise ValueError(f"could not read {image}")
        img = cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB)
    elif isinstance(image, bytes):
        nparr = np.frombuffer(image, np.uint8)
        img = cv2.imdecode(nparr, cv2.IMREAD_COLOR)
        img = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
        img_cv_grey = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
    elif isinstance(image, np.ndarray):
        if image.ndim == 2:
            img_cv_grey = image
            img = cv2.cvtColor(image, cv2.COLOR_GRAY2BGR)
        elif image.ndim == 3 and image.shape[2] == 1:
            img_cv_grey = np.squeeze(image)
            img = cv2.cvtColor(img_cv_grey, cv2.COLOR_GRAY2BGR)
        elif image.ndim == 3 and image.shape[2] == 3:
            img = image
            img_cv_grey = cv2.cvtColor(image, cv2.COLOR_BGR2GRAY)
        elif image.ndim == 3 and image.shape[2] == 4:
            img = cv2.cvtColor(image[:, :, :3], cv2.COLOR_RGB2BGR)
            img_cv_grey = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
        else:
            raise ValueError("Invalid input type")
    elif hasattr(image, "convert") and hasattr(image, "size"):          # PIL image (upstream: JpegImageFile)
        image_array = np.array(image.convert("RGB"))
        img = cv2.cvtColor(image_array, cv2.COLOR_RGB2BGR)
        img_cv_grey = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
    else:
        raise ValueError("Invalid input type. Supporting format = string(file path or url), bytes, numpy array")
    return img, img_cv_grey


def resize_aspect_ratio(img, square_size, interpolation, mag_ratio=1.0):
    height, width, channel = img.shape
    target_size = mag_ratio * max(height, width)
    if target_size > square_size:
        target_size = square_size
    ratio = target_size / max(height, width)
    target_h, target_w = int(height * ratio), int(width * ratio)
    proc = cv2.resize(img, (target_w, target_h), interpolation=interpolation)
    target_h32, target_w32 = target_h, target_w
    if target_h % 32 != 0:
        target_h32 = target_h + (32 - target_h % 32)
    if target_w % 32 != 0:
        target_w32 = target_w + (32 - target_w % 32)
    resized = np.zeros((target_h32, target_w32, channel), dtype=np.float32)
    resized[0:target_h, 0:target_w, :] = proc
    return resized, ratio, (int(target_w32 / 2), int(target_h32 / 2))


def normalize_mean_variance(in_img, mean=(0.485, 0.456, 0.406), variance=(0.229, 0.224, 0.225)):
    img = in_img.copy().astype(np.float32)
    img -= np.array([mean[0] * 255.0, mean[1] * 255.0, mean[2] * 255.0], dtype=np.float32)
    img /= np.array([variance[0] * 255.0, variance[1] * 255.0, variance[2] * 255.0], dtype=np.float32)
    return img


def canvas_tensor(img, canvas_size=2560, mag_ratio=1.0):
    """detection.test_net up to the network input.  -> (1x3xHxW float32 tensor, ratio)."""
    resized, ratio, _ = resize_aspect_ratio(img, canvas_size, cv2.INTER_LINEAR, mag_ratio)
    x = np.transpose(normalize_mean_variance(resized), (2, 0, 1))
    return torch.from_numpy(np.ascontiguousarray(x[None])), ratio


def get_det_boxes_core(textmap, linkmap, text_threshold, link_threshold, low_text):
    """craft_utils.getDetBoxes_core (estimate_num_chars=False).  -> list of 4x2 float32 boxes, labels, mapper."""
    linkmap = linkmap.copy()
    textmap = textmap.copy()
    img_h, img_w = textmap.shape
    _, text_score = cv2.threshold(textmap, low_text, 1, 0)
    _, link_score = cv2.threshold(linkmap, link_threshold, 1, 0)
    text_score_comb = np.clip(text_score + link_score, 0, 1)
    nLabels, labels, stats, _ = cv2.connectedComponentsWithStats(text_score_comb.astype(np.uint8), connectivity=4)
    det, mapper = [], []
    for k in range(1, nLabels):
        size = stats[k, cv2.CC_STAT_AREA]
        if size < 10:
            continue
        if np.max(textmap[labels == k]) < text_threshold:
            continue
        segmap = np.zeros(textmap.shape, dtype=np.uint8)
        segmap[labels == k] = 255
        mapper.append(k)
        segmap[np.logical_and(link_score == 1, text_score == 0)] = 0
        x, y = stats[k, cv2.CC_STAT_LEFT], stats[k, cv2.CC_STAT_TOP]
        w, h = stats[k, cv2.CC_STAT_WIDTH], stats[k, cv2.CC_STAT_HEIGHT]
        niter = int(math.sqrt(size * min(w, h) / (w * h)) * 2)
        sx, ex, sy, ey = x - niter, x + w + niter + 1, y - niter, y + h + niter + 1
        if sx < 0:
            sx = 0
        if sy < 0:
            sy = 0
        if ex >= img_w:
            ex = img_w
        if ey >= img_h:
            ey = img_h
        kernel = cv2.getStructuringElement(cv2.MORPH_RECT, (1 + niter, 1 + niter))
        segmap[sy:ey, sx:ex] = cv2.dilate(segmap[sy:ey, sx:ex], kernel)
        np_contours = np.roll(np.array(np.where(segmap != 0)), 1, axis=0).transpose().reshape(-1, 2)
        rectangle = cv2.minAreaRect(np_contours)
        box = cv2.boxPoints(rectangle)
        w, h = np.linalg.norm(box[0] - box[1]), np.linalg.norm(box[1] - box[2])
        box_ratio = max(w, h) / (min(w, h) + 1e-5)
        if abs(1 - box_ratio) <= 0.1:
            l, r = min(np_contours[:, 0]), max(np_contours[:, 0])
            t, b = min(np_contours[:, 1]), max(np_contours[:, 1])
            box = np.array([[l, t], [r, t], [r, b], [l, b]], dtype=np.float32)
        startidx = box.sum(axis=1).argmin()
        box = np.roll(box, 4 - startidx, 0)
        det.append(np.array(box))
    return det, labels, mapper


def adjust_result_coordinates(polys, ratio_w, ratio_h, ratio_net=2):
    if len(polys) > 0:
        polys = np.array(polys)
        for k in range(len(polys)):
            if polys[k] is not None:
                polys[k] *= (ratio_w * ratio_net, ratio_h * ratio_net)
    return polys


def boxes_to_polys_int(boxes, ratio):
    """detection.test_net tail + get_textbox: scale by 2/ratio, truncate to int32, flatten to 8 ints."""
    ratio_h = ratio_w = 1 / ratio
    boxes = adjust_result_coordinates(boxes, ratio_w, ratio_h)
    return [np.array(b).astype(np.int32).reshape((-1)) for b in boxes]


def group_text_box(polys, slope_ths=0.1, ycenter_ths=0.5, height_ths=0.5, width_ths=1.0, add_margin=0.05,
                   sort_output=True):
    """easyocr/utils.py::group_text_box"""
    horizontal_list, free_list, combined_list, merged_list = [], [], [], []
    for poly in polys:
        slope_up = (poly[3] - poly[1]) / np.maximum(10, (poly[2] - poly[0]))
        slope_down = (poly[5] - poly[7]) / np.maximum(10, (poly[4] - poly[6]))
        if max(abs(slope_up), abs(slope_down)) < slope_ths:
            x_max = max([poly[0], poly[2], poly[4], poly[6]])
            x_min = min([poly[0], poly[2], poly[4], poly[6]])
            y_max = max([poly[1], poly[3], poly[5], poly[7]])
            y_min = min([poly[1], poly[3], poly[5], poly[7]])
            horizontal_list.append([x_min, x_max, y_min, y_max, 0.5 * (y_min + y_max), y_max - y_min])
        else:
            height = np.linalg.norm([poly[6] - poly[0], poly[7] - poly[1]])
            width = np.linalg.norm([poly[2] - poly[0], poly[3] - poly[1]])
            margin = int(1.44 * add_margin * min(width, height))
            theta13 = abs(np.arctan((poly[1] - poly[5]) / np.maximum(10, (poly[0] - poly[4]))))
            theta24 = abs(np.arctan((poly[3] - poly[7]) / np.maximum(10, (poly[2] - poly[6]))))
            x1 = poly[0] - np.cos(theta13) * margin
            y1 = poly[1] - np.sin(theta13) * margin
            x2 = poly[2] + np.cos(theta24) * margin
            y2 = poly[3] - np.sin(theta24) * margin
            x3 = poly[4] + np.cos(theta13) * margin
            y3 = poly[5] + np.sin(theta13) * margin
            x4 = poly[6] - np.cos(theta24) * margin
            y4 = poly[7] + np.sin(theta24) * margin
            free_list.append([[x1, y1], [x2, y2], [x3, y3], [x4, y4]])
    if sort_output:
        horizontal_list = sorted(horizontal_list, key=lambda item: item[4])
    new_box = []
    for poly in horizontal_list:
        if len(new_box) == 0:
            b_height = [poly[5]]
            b_ycenter = [poly[4]]
            new_box.append(poly)
        else:
            if abs(np.mean(b_ycenter) - poly[4]) < ycenter_ths * np.mean(b_height):
                b_height.append(poly[5])
                b_ycenter.append(poly[4])
                new_box.append(poly)
            else:
                b_height = [poly[5]]
                b_ycenter = [poly[4]]
                combined_list.append(new_box)
                new_box = [poly]
    combined_list.append(new_box)
    for boxes in combined_list:
        if len(boxes) == 1:
            box = boxes[0]
            margin = int(add_margin * min(box[1] - box[0], box[5]))
            merged_list.append([box[0] - margin, box[1] + margin, box[2] - margin, box[3] + margin])
        else:
            boxes = sorted(boxes, key=lambda item: item[0])
            merged_box, new_box = [], []
            for box in boxes:
                if len(new_box) == 0:
                    b_height = [box[5]]
                    x_max = box[1]
                    new_box.append(box)
                else:
                    if (abs(np.mean(b_height) - box[5]) < height_ths * np.mean(b_height)) and \
                            ((box[0] - x_max) < width_ths * (box[3] - box[2])):
                        b_height.append(box[5])
                        x_max = box[1]
                        new_box.append(box)
                    else:
                        b_height = [box[5]]
                        x_max = box[1]
                        merged_box.append(new_box)
                        new_box = [box]
            if len(new_box) > 0:
                merged_box.append(new_box)
            for mbox in merged_box:
                if len(mbox) != 1:
                    x_min = min(mbox, key=lambda x: x[0])[0]
                    x_max = max(mbox, key=lambda x: x[1])[1]
                    y_min = min(mbox, key=lambda x: x[2])[2]
                    y_max = max(mbox, key=lambda x: x[3])[3]
                    box_width = x_max - x_min
                    box_height = y_max - y_min
                    margin = int(add_margin * (min(box_width, box_height)))
                    merged_list.append([x_min - margin, x_max + margin, y_min - margin, y_max + margin])
                else:
                    box = mbox[0]
                    box_width = box[1] - box[0]
                    box_height = box[3] - box[2]
                    margin = int(add_margin * (min(box_width, box_height)))
                    merged_list.append([box[0] - margin, box[1] + margin, box[2] - margin, box[3] + margin])
    return merged_list, free_list


def _diff(values):
    return max(values) - min(values)


def filter_min_size(horizontal_list, free_list, min_size=20):
    """Reader.detect tail."""
    if min_size:
        horizontal_list = [i for i in horizontal_list if max(i[1] - i[0], i[3] - i[2]) > min_size]
        free_list = [i for i in free_list if max(_diff([c[0] for c in i]), _diff([c[1] for c in i])) > min_size]
    return horizontal_list, free_list


# ======================================================================================================================
# Crops (easyocr/utils.py)
# ======================================================================================================================

def four_point_transform(image, rect):
    (tl, tr, br, bl) = rect
    widthA = np.sqrt(((br[0] - bl[0]) ** 2) + ((br[1] - bl[1]) ** 2))
    widthB = np.sqrt(((tr[0] - tl[0]) ** 2) + ((tr[1] - tl[1]) ** 2))
    maxWidth = max(int(widthA), int(widthB))
    heightA = np.sqrt(((tr[0] - br[0]) ** 2) + ((tr[1] - br[1]) ** 2))
    heightB = np.sqrt(((tl[0] - bl[0]) ** 2) + ((tl[1] - bl[1]) ** 2))
    maxHeight = max(int(heightA), int(heightB))
    dst = np.array([[0, 0], [maxWidth - 1, 0], [maxWidth - 1, maxHeight - 1], [0, maxHeight - 1]], dtype="float32")
    M = cv2.getPerspectiveTransform(rect, dst)
    return cv2.warpPerspective(image, M, (maxWidth, maxHeight))


def calculate_ratio(width, height):
    ratio = width / height
    if ratio < 1.0:
        ratio = 1.0 / ratio
    return ratio


def compute_ratio_and_resize(img, width, height, model_height):
    """upstream passes interpolation=Image.Resampling.LANCZOS (== 1 == cv2.INTER_LINEAR) to cv2.resize."""
    ratio = width / height
    if ratio < 1.0:
        ratio = calculate_ratio(width, height)
        img = cv2.resize(img, (model_height, int(model_height * ratio)), interpolation=cv2.INTER_LINEAR)
    else:
        img = cv2.resize(img, (int(model_height * ratio), model_height), interpolation=cv2.INTER_LINEAR)
    return img, ratio


def get_image_list(horizontal_list, free_list, img, model_height=64, sort_output=True):
    image_list = []
    maximum_y, maximum_x = img.shape
    max_ratio_hori, max_ratio_free = 1, 1
    for box in free_list:
        rect = np.array(box, dtype="float32")
        transformed_img = four_point_transform(img, rect)
        ratio = calculate_ratio(transformed_img.shape[1], transformed_img.shape[0])
        new_width = int(model_height * ratio)
        if new_width == 0:
            pass
        else:
            crop_img, ratio = compute_ratio_and_resize(transformed_img, transformed_img.shape[1],
                                                       transformed_img.shape[0], model_height)
            image_list.append((box, crop_img))
            max_ratio_free = max(ratio, max_ratio_free)
    max_ratio_free = math.ceil(max_ratio_free)
    for box in horizontal_list:
        x_min = max(0, box[0])
        x_max = min(box[1], maximum_x)
        y_min = max(0, box[2])
        y_max = min(box[3], maximum_y)
        crop_img = img[y_min:y_max, x_min:x_max]
        width = x_max - x_min
        height = y_max - y_min
        ratio = calculate_ratio(width, height)
        new_width = int(model_height * ratio)
        if new_width == 0:
            pass
        else:
            crop_img, ratio = compute_ratio_and_resize(crop_img, width, height, model_height)
            image_list.append(([[x_min, y_min], [x_max, y_min], [x_max, y_max], [x_min, y_max]], crop_img))
            max_ratio_hori = max(ratio, max_ratio_hori)
    max_ratio_hori = math.ceil(max_ratio_hori)
    max_ratio = max(max_ratio_hori, max_ratio_free)
    max_width = math.ceil(max_ratio) * model_height
    if sort_output:
        image_list = sorted(image_list, key=lambda item: item[0][0][1])
    return image_list, max_width


# ======================================================================================================================
# Recognition (easyocr/recognition.py, utils.CTCLabelConverter)
# ======================================================================================================================

def contrast_grey(img):
    high = np.percentile(img, 90)
    low = np.percentile(img, 10)
    return (high - low) / np.maximum(10, high + low), high, low


def adjust_contrast_grey(img, target=0.4):
    contrast, high, low = contrast_grey(img)
    if contrast < target:
        img = img.astype(int)
        ratio = 200.0 / np.maximum(10, high - low)
        img = (img - low + 25) * ratio
        img = np.maximum(np.full(img.shape, 0), np.minimum(np.full(img.shape, 255), img)).astype(np.uint8)
    return img


def align_collate_one(crop_u8, imgW, adjust_contrast=0.0):
    """AlignCollate(imgH=64, imgW, keep_ratio_with_pad=True) + NormalizePAD for one crop -> float32 (64, imgW)."""
    image = Image.fromarray(crop_u8, "L")
    w, h = image.size
    if adjust_contrast > 0:
        arr = np.array(image.convert("L"))
        arr = adjust_contrast_grey(arr, target=adjust_contrast)
        image = Image.fromarray(arr, "L")
    ratio = w / float(h)
    if math.ceil(IMG_H * ratio) > imgW:
        resized_w = imgW
    else:
        resized_w = math.ceil(IMG_H * ratio)
    resized = image.resize((resized_w, IMG_H), Image.BICUBIC)
    a = np.asarray(resized).astype(np.float32) / np.float32(255.0)      # ToTensor
    a = (a - np.float32(0.5)) / np.float32(0.5)
    out = np.zeros((IMG_H, imgW), np.float32)
    out[:, :resized_w] = a
    if imgW != resized_w:
        out[:, resized_w:] = a[:, resized_w - 1][:, None]
    return out


def custom_mean(x):
    return x.prod() ** (2.0 / np.sqrt(len(x)))


def decode_greedy(text_index, length, character=CHARACTERS):
    """CTCLabelConverter.decode_greedy; converter.character = ['[blank]'] + list(character); ignore_idx = [0]."""
    table = np.array(["[blank]"] + list(character))
    texts, index = [], 0
    for l in length:
        t = text_index[index:index + l]
        a = np.insert(~((t[1:] == t[:-1])), 0, True)
        b = ~np.isin(t, np.array([0]))
        c = a & b
        texts.append("".join(table[t[c.nonzero()]]))
        index += l
    return texts


def probs_from_logits(preds: torch.Tensor, ignore_idx=()):
    """recognizer_predict: softmax, zero ignored classes, renormalise.  preds (B,T,C) f32 -> numpy f32 (B,T,C)."""
    p = F.softmax(preds, dim=2).cpu().detach().numpy()
    p[:, :, list(ignore_idx)] = 0.0
    norm = p.sum(axis=2)
    return (p / np.expand_dims(norm, axis=-1)).astype(np.float32)


def decode_probs(preds_prob: np.ndarray, character=CHARACTERS):
    """Greedy branch of recognizer_predict.  -> list of [text, confidence]."""
    B, T, _ = preds_prob.shape
    idx = torch.from_numpy(preds_prob).float().max(2)[1].view(-1).numpy()
    strs = decode_greedy(idx, [T] * B, character)
    values = preds_prob.max(axis=2)
    indices = preds_prob.argmax(axis=2)
    out = []
    for s, v, i in zip(strs, values, indices):
        mp = v[i != 0]
        if len(mp) == 0:
            mp = np.array([0])
        out.append([s, custom_mean(mp)])
    return out


# ----------------------------------------------------------------------------------------------------------------------
# Beam-search decoders (easyocr/utils.py: BeamEntry, BeamState, fast_simplify_label, ctcBeamSearch,
# CTCLabelConverter.decode_beamsearch / decode_wordbeamsearch).  Restated from the 1.7.x source; the reference pins
# numpy==1.26.4, under which `python_float * np.float32` promotes to float64, so every beam probability is a float64
# product of float32 inputs -- written out explicitly here so that NumPy 2's scalar rules cannot change it.
# ----------------------------------------------------------------------------------------------------------------------

class _BeamEntry:
    __slots__ = ("prTotal", "prNonBlank", "prBlank", "prText", "labeling")

    def __init__(self):
        self.prTotal = 0.0
        self.prNonBlank = 0.0
        self.prBlank = 0.0
        self.prText = 1.0          # LM score; no language model is ever applied upstream (applyLM is commented out)
        self.labeling = ()


def fast_simplify_label(labeling, c, blankIdx=0):
    if labeling and c == blankIdx and labeling[-1] != blankIdx:          # blank after a character: keep it
        return labeling + (c,)
    if labeling and c != blankIdx and labeling[-1] == blankIdx:          # character after a blank
        if labeling[-2] == c:                                            # blank between equal characters stays
            return labeling + (c,)
        return labeling[:-1] + (c,)                                      # blank between different characters goes
    if labeling and c == blankIdx and labeling[-1] == blankIdx:          # consecutive blanks collapse
        return labeling
    if not labeling and c == blankIdx:                                   # leading blank is dropped
        return labeling
    return labeling + (c,)


def _sorted_beams(entries):
    """BeamState.sort: sorted(..., reverse=True, key=prTotal*prText) -- stable, ties keep insertion order."""
    return sorted(entries.values(), reverse=True, key=lambda x: x.prTotal * x.prText)


def ctc_beam_search(mat, character, ignore_idx, beamWidth=25, dict_list=()):
    """utils.ctcBeamSearch(mat, classes, ignore_idx, lm=None, beamWidth, dict_list).  mat: (T, C) float32 probabilities;
    `character` = converter.character (['[blank]'] + alphabet)."""
    blankIdx = 0
    maxT, maxC = mat.shape
    m = [[float(v) for v in row] for row in mat]                        # float32 -> float64, exact
    last = {(): _BeamEntry()}
    last[()].prBlank = 1.0
    last[()].prTotal = 1.0
    for t in range(maxT):
        curr = {}
        best = [b.labeling for b in _sorted_beams(last)[0:beamWidth]]
        # the comparison `float32_array >= python_float` is made in float32 under NumPy 1.26 (value-based casting of the
        # scalar) as well as under NumPy 2 (weak scalar): round the threshold to float32 first
        thr32 = float(np.float32(0.5 / maxC))
        cand = [c for c in range(maxC) if m[t][c] >= thr32]
        for labeling in best:
            prNonBlank = 0.0
            if labeling:
                prNonBlank = last[labeling].prNonBlank * m[t][labeling[-1]]
            prBlank = last[labeling].prTotal * m[t][blankIdx]
            e = curr.get(labeling)
            if e is None:
                e = curr[labeling] = _BeamEntry()
            e.labeling = labeling
            e.prNonBlank += prNonBlank
            e.prBlank += prBlank
            e.prTotal += prBlank + prNonBlank
            e.prText = last[labeling].prText
            for c in cand:
                newLabeling = fast_simplify_label(labeling, c, blankIdx)
                if labeling and labeling[-1] == c:
                    prNonBlank = m[t][c] * last[labeling].prBlank
                else:
                    prNonBlank = m[t][c] * last[labeling].prTotal
                e2 = curr.get(newLabeling)
                if e2 is None:
                    e2 = curr[newLabeling] = _BeamEntry()
                e2.labeling = newLabeling
                e2.prNonBlank += prNonBlank
                e2.prTotal += prNonBlank
        last = curr
    for e in last.values():                                              # BeamState.norm (prText stays 1.0 without an LM)
        n = len(e.labeling)
        e.prText = e.prText ** (1.0 / (n if n else 1.0))

    def text_of(lab):
        return "".join(character[l] for i, l in enumerate(lab) if l not in ignore_idx and not (i > 0 and lab[i - 1] == lab[i]))

    beams = _sorted_beams(last)
    if len(dict_list) == 0:
        return text_of(beams[0].labeling)
    best_text = None                                                     # BeamState.wordsearch(classes, ignore_idx, 20, dict_list)
    for j, cand_beam in enumerate(beams[:20]):
        text = text_of(cand_beam.labeling)
        if j == 0:
            best_text = text
        if text in dict_list:
            best_text = text
            break
    return best_text


def decode_beamsearch(preds_prob, character=CHARACTERS, beamWidth=5):
    table = ["[blank]"] + list(character)
    return [ctc_beam_search(preds_prob[i], table, [0], beamWidth=beamWidth) for i in range(preds_prob.shape[0])]


def decode_wordbeamsearch(preds_prob, character=CHARACTERS, beamWidth=5, dict_list=()):
    """CTCLabelConverter.decode_wordbeamsearch, the branch without separator characters (every Latin-script model):
    the arg-max path is cut at its space symbols and every run in between is beam-searched against the dictionary."""
    table = ["[blank]"] + list(character)
    space_idx = character.index(" ") + 1
    argmax = np.argmax(preds_prob, axis=2)
    texts = []
    for i in range(preds_prob.shape[0]):
        string = ""
        data = np.argwhere(argmax[i] != space_idx).flatten()
        group = np.split(data, np.where(np.diff(data) != 1)[0] + 1)
        group = [list(item) for item in group if len(item) > 0]
        for j, list_idx in enumerate(group):
            t = ctc_beam_search(preds_prob[i, list_idx, :], table, [0], beamWidth=beamWidth, dict_list=dict_list)
            string += t if j == 0 else " " + t
        texts.append(string)
    return texts


def make_rotated_img_list(rotation_info, img_list):
    """utils.make_rotated_img_list: scipy.ndimage.rotate(img, angle, reshape=True) of every crop for every angle.  For the
    documented angles (90, 180, 270) the spline rotation returns exactly np.rot90(img, angle // 90) (checked against scipy
    1.18 on 600 random u8 crops, tests/test_oracle_easyocr.py) -- restated that way so the oracle does not need scipy."""
    out = list(img_list)
    for angle in rotation_info:
        if angle not in (90, 180, 270):
            raise ValueError("rotation_info: eligible values are 90, 180 and 270")
        for box, img in img_list:
            out.append((box, np.ascontiguousarray(np.rot90(img, angle // 90))))
    return out


def set_result_with_confidence(results):
    final = []
    for col in range(len(results[0])):
        best_row = max([(row, results[row][col][2]) for row in range(len(results))], key=lambda x: x[1])[0]
        final.append(results[best_row][col])
    return final


# ======================================================================================================================
# Reader
# ======================================================================================================================

def get_paragraph(raw_result, x_ths=1, y_ths=0.5, mode="ltr"):
    """easyocr/utils.py::get_paragraph, restated statement by statement (rows are
    [text, min_x, max_x, min_y, max_y, height, y_centre, group])."""
    box_group = []
    for box in raw_result:
        all_x = [int(coord[0]) for coord in box[0]]
        all_y = [int(coord[1]) for coord in box[0]]
        min_x, max_x, min_y, max_y = min(all_x), max(all_x), min(all_y), max(all_y)
        box_group.append([box[1], min_x, max_x, min_y, max_y, max_y - min_y, 0.5 * (min_y + max_y), 0])
    current_group = 1
    while len([b for b in box_group if b[7] == 0]) > 0:
        box_group0 = [b for b in box_group if b[7] == 0]
        if len([b for b in box_group if b[7] == current_group]) == 0:
            box_group0[0][7] = current_group
        else:
            cur = [b for b in box_group if b[7] == current_group]
            mean_height = np.mean([b[5] for b in cur])
            min_gx = min(b[1] for b in cur) - x_ths * mean_height
            max_gx = max(b[2] for b in cur) + x_ths * mean_height
            min_gy = min(b[3] for b in cur) - y_ths * mean_height
            max_gy = max(b[4] for b in cur) + y_ths * mean_height
            add_box = False
            for b in box_group0:
                same_horizontal_level = (min_gx <= b[1] <= max_gx) or (min_gx <= b[2] <= max_gx)
                same_vertical_level = (min_gy <= b[3] <= max_gy) or (min_gy <= b[4] <= max_gy)
                if same_horizontal_level and same_vertical_level:
                    b[7] = current_group
                    add_box = True
                    break
            if not add_box:
                current_group += 1
    result = []
    for i in set(b[7] for b in box_group):
        cur = [b for b in box_group if b[7] == i]
        mean_height = np.mean([b[5] for b in cur])
        min_gx, max_gx = min(b[1] for b in cur), max(b[2] for b in cur)
        min_gy, max_gy = min(b[3] for b in cur), max(b[4] for b in cur)
        text = ""
        while len(cur) > 0:
            highest = min(b[6] for b in cur)
            candidates = [b for b in cur if b[6] < highest + 0.4 * mean_height]
            if mode == "ltr":
                most_left = min(b[1] for b in candidates)
                for b in candidates:
                    if b[1] == most_left:
                        best_box = b
            elif mode == "rtl":
                most_right = max(b[2] for b in candidates)
                for b in candidates:
                    if b[2] == most_right:
                        best_box = b
            text += " " + best_box[0]
            cur.remove(best_box)
        result.append([[[min_gx, min_gy], [max_gx, min_gy], [max_gx, max_gy], [min_gx, max_gy]], text[1:]])
    return result


class Reader:
    """Restated easyocr.Reader(['en']) -- CPU, greedy decoder, the options BB-OCR exercises."""

    def __init__(self, craft: CRAFT, crnn: CRNN, quantize: bool = False, character: str = CHARACTERS):
        self.detector = craft.eval()
        self.recognizer = crnn.eval()
        if quantize:     # easyocr.recognition.get_recognizer on CPU with quantize=True
            self.recognizer = torch.quantization.quantize_dynamic(self.recognizer, dtype=torch.qint8)
        self.character = character
        self.ignore_idx = []       # character - lang_char is empty for ['en'] + english_g2
        self.dict_list = []        # easyocr/dict/en.txt (not in this image); wordbeamsearch without it = per-word beam search

    # ---- detection -----------------------------------------------------------------------------------------------
    def score_maps(self, img, canvas_size=2560, mag_ratio=1.0):
        x, ratio = canvas_tensor(img, canvas_size, mag_ratio)
        with torch.no_grad():
            y, _ = self.detector(x)
        return y[0, :, :, 0].numpy().copy(), y[0, :, :, 1].numpy().copy(), ratio

    def boxes_from_maps(self, score_text, score_link, ratio, min_size=20, text_threshold=0.7, low_text=0.4,
                        link_threshold=0.4, slope_ths=0.1, ycenter_ths=0.5, height_ths=0.5, width_ths=0.5,
                        add_margin=0.1):
        boxes, _, _ = get_det_boxes_core(score_text, score_link, text_threshold, link_threshold, low_text)
        polys = boxes_to_polys_int(boxes, ratio)
        h_list, f_list = group_text_box(polys, slope_ths, ycenter_ths, height_ths, width_ths, add_margin, True)
        return filter_min_size(h_list, f_list, min_size)

    def detect(self, img, **kw):
        canvas_size = kw.pop("canvas_size", 2560)
        mag_ratio = kw.pop("mag_ratio", 1.0)
        st, sl, ratio = self.score_maps(img, canvas_size, mag_ratio)
        return self.boxes_from_maps(st, sl, ratio, **kw)

    # ---- recognition ---------------------------------------------------------------------------------------------
    def logits(self, batch: np.ndarray):
        """batch: (B, 64, W) float32 normalised -> (B, T, 97) float32 logits."""
        with torch.no_grad():
            return self.recognizer(torch.from_numpy(batch[:, None]), None)

    def _predict(self, crops, imgW, adjust_contrast=0.0, decoder="greedy", beamWidth=5):
        res = []
        for c in crops:                                     # one crop per forward pass (the result does not depend on the batch)
            x = align_collate_one(c, imgW, adjust_contrast)[None]
            probs = probs_from_logits(self.logits(x), self.ignore_idx)
            one = decode_probs(probs, self.character)       # greedy string + the confidence every decoder reports
            if decoder == "beamsearch":
                one[0][0] = decode_beamsearch(probs, self.character, beamWidth)[0]
            elif decoder == "wordbeamsearch":
                one[0][0] = decode_wordbeamsearch(probs, self.character, beamWidth, self.dict_list)[0]
            res += one
        return res

    def get_text(self, image_list, imgW, contrast_ths=0.1, adjust_contrast=0.5, decoder="greedy", beamWidth=5):
        coord = [item[0] for item in image_list]
        img_list = [item[1] for item in image_list]
        result1 = self._predict(img_list, imgW, 0.0, decoder, beamWidth)
        low = [i for i, item in enumerate(result1) if item[1] < contrast_ths]
        result2 = self._predict([img_list[i] for i in low], imgW, adjust_contrast, decoder, beamWidth) if low else []
        result = []
        for i, (box, pred1) in enumerate(zip(coord, result1)):
            if i in low:
                pred2 = result2[low.index(i)]
                if pred1[1] > pred2[1]:
                    result.append((box, pred1[0], pred1[1]))
                else:
                    result.append((box, pred2[0], pred2[1]))
            else:
                result.append((box, pred1[0], pred1[1]))
        return result

    def recognize(self, img_cv_grey, horizontal_list, free_list, contrast_ths=0.1, adjust_contrast=0.5, decoder="greedy",
                  beamWidth=5, batch_size=1, rotation_info=None):
        """Reader.recognize: the per-box branch (batch_size == 1 and no rotation_info; what BB-OCR runs) or upstream's
        batched branch (one max_width for all boxes of the page, crops ordered by y, optional rotated copies)."""
        if batch_size == 1 and not rotation_info:
            result = []
            for bbox in horizontal_list:
                image_list, max_width = get_image_list([bbox], [], img_cv_grey, model_height=IMG_H)
                result += self.get_text(image_list, int(max_width), contrast_ths, adjust_contrast, decoder, beamWidth)
            for bbox in free_list:
                image_list, max_width = get_image_list([], [bbox], img_cv_grey, model_height=IMG_H)
                result += self.get_text(image_list, int(max_width), contrast_ths, adjust_contrast, decoder, beamWidth)
            return result
        image_list, max_width = get_image_list(horizontal_list, free_list, img_cv_grey, model_height=IMG_H)
        image_len = len(image_list)
        if rotation_info and image_list:
            image_list = make_rotated_img_list(rotation_info, image_list)
            max_width = max(max_width, IMG_H)
        result = self.get_text(image_list, int(max_width), contrast_ths, adjust_contrast, decoder, beamWidth)
        if rotation_info and (horizontal_list + free_list):
            result = set_result_with_confidence([result[image_len * i:image_len * (i + 1)] for i in range(len(rotation_info) + 1)])
        return result

    def readtext(self, image, min_size=20, contrast_ths=0.1, adjust_contrast=0.5, text_threshold=0.7, low_text=0.4,
                 link_threshold=0.4, canvas_size=2560, mag_ratio=1.0, slope_ths=0.1, ycenter_ths=0.5, height_ths=0.5,
                 width_ths=0.5, add_margin=0.1, decoder="greedy", beamWidth=5, batch_size=1, rotation_info=None, **_ignored):
        img, img_cv_grey = reformat_input(image)
        h_list, f_list = self.detect(img, min_size=min_size, text_threshold=text_threshold, low_text=low_text,
                                     link_threshold=link_threshold, canvas_size=canvas_size, mag_ratio=mag_ratio,
                                     slope_ths=slope_ths, ycenter_ths=ycenter_ths, height_ths=height_ths,
                                     width_ths=width_ths, add_margin=add_margin)
        return self.recognize(img_cv_grey, h_list, f_list, contrast_ths, adjust_contrast, decoder, beamWidth, batch_size,
                              rotation_info)
