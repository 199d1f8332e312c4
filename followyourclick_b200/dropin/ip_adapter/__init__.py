"""Drop-in shim package: modules defined here shadow the reference's, everything else falls through to the
reference tree that follows this directory on sys.path (pkgutil.extend_path).  See INTEGRATION.md."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
from followyourclick_b200.ip_adapter import (IPAttnProcessor, IPAttnProcessor2_0, MyIPAdapter, MyIPAdapterPlus,  # noqa: E402,F401
                                             Resampler)
