"""Short import alias for the hot-path package.

The package proper lives in the directory the project contract names
(``automated-brain-mri-analysis-and-report-generation-with-retrieval-augmented-clinical-assistance_b200/``); that
name is not a Python identifier, so this shim makes ``import brainseg_b200.<module>`` resolve to the files there.
"""
import os as _os

_PKG_DIR = _os.path.join(
    _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
    "automated-brain-mri-analysis-and-report-generation-with-retrieval-augmented-clinical-assistance_b200",
)
__path__ = [_PKG_DIR]
PACKAGE_DIR = _PKG_DIR
with open(_os.path.join(_PKG_DIR, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_PKG_DIR, "__init__.py"), "exec"))
