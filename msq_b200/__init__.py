"""Importable alias of the `microscopiq-llm-quantization_b200/` package directory (whose
mandated name is not a valid Python identifier): `import msq_b200`, `msq_b200.utils.quant`..."""
import os as _os

_impl = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "microscopiq-llm-quantization_b200")
__path__.insert(0, _impl)
with open(_os.path.join(_impl, "__init__.py")) as _fh:
    exec(compile(_fh.read(), _os.path.join(_impl, "__init__.py"), "exec"))
