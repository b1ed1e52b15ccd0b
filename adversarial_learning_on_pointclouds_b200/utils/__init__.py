from .model_utils import init_net, init_weights, load_models           # noqa: F401
from .utils import make_D_label, make_shape_label                     # noqa: F401
from .image_pool import ImagePool                                     # noqa: F401
from .metric import batch_get_iou, get_iou, part_iou_from_logits, object_names, seg_classes   # noqa: F401
