from . import llama_cpp  # noqa: F401  (registration side effect)
