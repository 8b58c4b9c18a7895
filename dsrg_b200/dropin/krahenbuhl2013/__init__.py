from .CRF import *  # noqa: F401,F403  (same re-export as the reference's CRF/krahenbuhl2013/__init__.py:1)
