"""Access to the UNMODIFIED reference module staged under baseline/_ref/ by __graft_entry__.build()
(baseline/_ref is git-ignored; it travels to the GPU box with the gpurun snapshot).  Used by bench.py's
reference / torch_gpu arms and by tests -- never by the product path."""
import importlib.util
import os

REF_FILE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "model", "bilinear.py")


def available():
    return os.path.exists(REF_FILE)


def load_reference_module():
    """The reference's model/bilinear.py as a module object (it imports nothing but torch)."""
    if not available():
        raise RuntimeError("baseline/_ref/model/bilinear.py is missing: run __graft_entry__.build() in the build "
                           "container (it copies the file from /root/reference)")
    spec = importlib.util.spec_from_file_location("reference_model_bilinear", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
