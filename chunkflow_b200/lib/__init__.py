"""Small host-side helpers mirrored from the reference's ``chunkflow.lib``."""
import importlib.util


def load_source(fname: str, name: str = "Model"):
    """Import a python file as a module (reference: chunkflow/lib/__init__.py:5-16)."""
    spec = importlib.util.spec_from_file_location(name, fname)
    if spec is None or spec.loader is None:
        raise ImportError(f"cannot load python source: {fname}")
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    return module
