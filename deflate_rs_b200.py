"""Import alias: the package directory is named `deflate-rs_b200/` (not a Python identifier), so this
module turns itself into that package.  `import deflate_rs_b200` == the code in deflate-rs_b200/."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "deflate-rs_b200")]
__package__ = __name__
if __spec__ is not None:
    __spec__.submodule_search_locations = __path__
__file__ = _os.path.join(__path__[0], "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
