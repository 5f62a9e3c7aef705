"""neuralrgbd_b200: B200-native plane-sweep depth-probability-volume engine behind the
NVlabs/neuralrgbd call surface (`models.KVNET.KVNET`, `warping.homography.*`, `mutils.misc`).

`install_as_reference_modules()` registers the mirrors under the reference's import names so
that an unmodified `test_utils/test_KVNet.py` (which does `import warping.homography as
warp_homo`, `import mutils.misc as m_misc`) runs on this engine.
"""
import sys


def install_as_reference_modules():
    import types
    from .warping import homography
    from .models import KVNET as kvnet_mod
    from .mutils import misc
    from . import warping as warping_pkg, models as models_pkg, mutils as mutils_pkg
    for name, mod in (('warping', warping_pkg), ('warping.homography', homography), ('models', models_pkg),
                      ('models.KVNET', kvnet_mod), ('mutils', mutils_pkg), ('mutils.misc', misc)):
        sys.modules[name] = mod
    return types.SimpleNamespace(homography=homography, KVNET=kvnet_mod, misc=misc)
