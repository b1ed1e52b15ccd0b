from .model_utils import init_net, init_weights, load_models           # noqa: F401
from .utils import make_D_label, make_shape_label                     # noqa: F401
from .image_pool import ImagePool                                     # noqa: F401
