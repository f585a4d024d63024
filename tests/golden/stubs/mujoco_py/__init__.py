"""Names the reference imports from mujoco_py (ZoneEnvBase.py:6, wrappers.py:4);
none is used on the step/reset/obs path."""


class MujocoException(Exception):
    pass


class MjViewer:
    pass


class MjRenderContextOffscreen:
    pass


class const:
    GEOM_SPHERE = 2
    GEOM_LABEL = 101
