"""Drop-in shim: put this directory in front of the reference checkout on PYTHONPATH and the reference's own
`sample_ldm.py` / `sample_ddpm.py` import the B200 implementation under the module names they already use."""
from ldm_image_generator_b200.modules import *  # noqa: F401,F403
