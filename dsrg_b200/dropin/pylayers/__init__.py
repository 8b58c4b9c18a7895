from .pylayers import *  # noqa: F401,F403  (pylayers/pylayers/__init__.py:1 of the reference does the same)
