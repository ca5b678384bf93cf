"""Make the reference's orchestrators run on ffwm_b200 without editing them.

`install()` registers this package's modules under the import names the reference uses
(`models.external_function`, `models.base_networks`, `models.losses`, `lightcnn.light_cnn`, and the
three pybind module names), so that `models/ffwm_model.py` / `models/flownet_model.py`
(`from . import losses, external_function, base_networks`, models/ffwm_model.py:3) bind to the
sm_100a kernels when imported afterwards from a reference checkout on sys.path.
"""
import sys
import types

from . import base_networks, dropin, external_function, light_cnn, losses


def install(package="models"):
    dropin.install()
    pkg = sys.modules.get(package)
    for name, mod in (("external_function", external_function), ("base_networks", base_networks), ("losses", losses)):
        sys.modules["%s.%s" % (package, name)] = mod
        if pkg is not None:
            setattr(pkg, name, mod)
    lc = sys.modules.get("lightcnn") or types.ModuleType("lightcnn")
    lc.light_cnn = light_cnn
    sys.modules.setdefault("lightcnn", lc)
    sys.modules["lightcnn.light_cnn"] = light_cnn
