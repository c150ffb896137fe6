"""Module path kept for reference compatibility (lib/layer_utils/nms/pth_nms.py:48-63)."""
from lib.layer_utils.nms_wrapper import nms as pth_nms  # noqa: F401
