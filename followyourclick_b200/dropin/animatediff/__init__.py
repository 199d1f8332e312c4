"""Drop-in shim package: modules defined here shadow the reference's, everything else falls through to the
reference tree that follows this directory on sys.path (pkgutil.extend_path).  See INTEGRATION.md."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
