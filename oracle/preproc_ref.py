"""CPU restatement of the reference's inference-time image pre-processing -- TEST INFRASTRUCTURE.

TaskPrompter/inference.py:127-133 (`infer_one_image`): cv2.imread (uint8 BGR, HWC) -> float32 -> BGR2RGB ->
`get_infer_transforms` (:93-115) = Normalize(mean, std) (data/transforms.py:236-251: x / 255, - mean, / std, all
float32) -> DirectResize (:66-81: cv2.resize INTER_LINEAR to the model's input size) -> ToTensor
(data/transforms.py:265-273: HWC -> CHW float) -> unsqueeze(0). InvPT/inference.py is the same pipeline.

cv2's float INTER_LINEAR samples at half-pixel centres with the source index clamped to the image (no
anti-aliasing), i.e. the same arithmetic as ATen's bilinear with align_corners=False; checked against cv2 itself in
tests/test_oracle.py::test_preproc_oracle_vs_golden (fixture made by oracle/make_golden.py with cv2.resize and the
reference's own Normalize / ToTensor classes)."""
import numpy as np

MEAN = (0.485, 0.456, 0.406)   # inference.py:99,107
STD = (0.229, 0.224, 0.225)


def infer_transform(img_bgr_u8, out_hw, mean=MEAN, std=STD):
    """img_bgr_u8: uint8 [h, w, 3] as cv2.imread returns it. Returns float32 [1, 3, H, W]."""
    img = img_bgr_u8.astype(np.float32)[:, :, ::-1]                      # inference.py:127-128 (BGR -> RGB)
    x = img / np.float32(255.0)                                          # transforms.py:248
    x = x - np.asarray(mean, np.float32).reshape(1, 1, 3)                # :249
    x = x / np.asarray(std, np.float32).reshape(1, 1, 3)                 # :250
    h, w = x.shape[:2]
    H, W = out_hw

    def coords(n_out, n_in):
        # cv2 (resize.cpp, INTER_LINEAR): fx = (dx + 0.5) * scale - 0.5 in double, sx = floor(fx), weight cast to float
        s = (np.arange(n_out, dtype=np.float64) + 0.5) * (n_in / n_out) - 0.5
        s = np.maximum(s, 0.0)
        i0 = np.minimum(np.floor(s).astype(np.int64), n_in - 1)
        i1 = np.minimum(i0 + 1, n_in - 1)
        return i0, i1, (s - i0).astype(np.float32)

    y0, y1, ly = coords(H, h)
    x0, x1, lx = coords(W, w)
    lx = lx.reshape(1, W, 1)
    ly = ly.reshape(H, 1, 1)
    top = x[y0][:, x0] * (1 - lx) + x[y0][:, x1] * lx                    # DirectResize, inference.py:77
    bot = x[y1][:, x0] * (1 - lx) + x[y1][:, x1] * lx
    out = top * (1 - ly) + bot * ly
    return np.ascontiguousarray(out.transpose(2, 0, 1))[None].astype(np.float32)   # ToTensor + unsqueeze(0)
