"""Import the UNMODIFIED reference (ishine/chunkformer) from /root/reference on CPU.

Only used by tests/golden/make_golden*.py (fixture generation in the build container).
Nothing that runs on the GPU box imports this module: /root/reference does not exist there
(bench.py's reference arm uses the copy installed under baseline/_ref through the same loader,
baseline/reference_arm.py).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from baseline import reference_arm as _ra  # noqa: E402

REFERENCE_ROOT = os.environ.get("CHUNKFORMER_REFERENCE", "/root/reference")
reference_config_dict = _ra.reference_config_dict


def import_reference():
    return _ra.import_reference(REFERENCE_ROOT)
