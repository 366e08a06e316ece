from .mlp import Mlp  # noqa: F401
from .drop import DropPath  # noqa: F401
from .patch_embed import PatchEmbed  # noqa: F401
