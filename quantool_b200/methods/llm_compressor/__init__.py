from . import awq, gptq, smoothquant  # noqa: F401  (registration side effect)
