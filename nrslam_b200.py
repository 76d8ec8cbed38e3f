"""Import alias: the package directory is `nr-slam_b200/` (not a valid Python identifier), exposed as
`nrslam_b200`."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "nr-slam_b200")]
__package__ = "nrslam_b200"
if __spec__ is not None:
    __spec__.submodule_search_locations = __path__
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
