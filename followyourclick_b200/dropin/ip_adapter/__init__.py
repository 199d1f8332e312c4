"""Drop-in shim package: modules defined here shadow the reference's, everything else falls through to the
reference tree that follows this directory on sys.path (pkgutil.extend_path).  See INTEGRATION.md."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
from followyourclick_b200.ip_adapter import IPAttnProcessor, IPAttnProcessor2_0, MyIPAdapter  # noqa: E402,F401

MyIPAdapterPlus = MyIPAdapter   # the Plus variant differs only in the projector (Perceiver Resampler: SURVEY 8f row 2)
