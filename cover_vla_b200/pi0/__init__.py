from .configuration_pi0 import PI0Config, PolicyFeature  # noqa: F401
from .modeling_pi0 import PI0FlowMatching, PI0Policy  # noqa: F401
