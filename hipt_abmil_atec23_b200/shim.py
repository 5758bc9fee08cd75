"""Helpers of the import-path shims (HIPT_4K/, models/, utils/, datasets/ at the repository root).

The reference's packages are directories WITHOUT __init__.py (namespace packages), so as soon as this repository is on
sys.path its same-named REGULAR shim packages win the import — whichever comes first.  A shim therefore has to keep every
other module of the reference's package reachable (`utils.file_utils`, `utils.eval_utils`, `models.resnet_custom`,
`datasets.dataset_h5`, `HIPT_4K.hipt_heatmap_utils` ...): `extend_package_path` appends the reference checkout's directory
of the same name to the shim package's __path__, and `chain_load` lets a shim MODULE (utils/utils.py) start from the
reference's module of the same name and override only what it accelerates.  With no reference checkout on sys.path both are
no-ops and the shims stand alone.
"""
import importlib.machinery
import os
import sys


def extend_package_path(pkg_name, pkg_path):
    """Append every `<sys.path entry>/<pkg_name>` directory that is a namespace portion (no __init__.py: a regular package
    such as site-packages' HuggingFace `datasets` is somebody else's) to the shim package's __path__."""
    own = {os.path.realpath(p) for p in pkg_path}
    for entry in list(sys.path):
        d = os.path.join(entry or os.getcwd(), pkg_name)
        real = os.path.realpath(d)
        if real in own or not os.path.isdir(d) or os.path.isfile(os.path.join(d, "__init__.py")):
            continue
        pkg_path.append(d)
        own.add(real)
    return pkg_path


def chain_load(module_globals, pkg_path, mod_name):
    """Execute the next `<mod_name>.py` (or sourceless `.pyc`) found on the package path BEYOND the shim's own directory into
    the shim module's namespace.  Returns the file it ran, or None when the shim stands alone."""
    full = module_globals["__name__"]
    for d in list(pkg_path)[1:]:
        for ext, loader_cls in ((".py", importlib.machinery.SourceFileLoader), (".pyc", importlib.machinery.SourcelessFileLoader)):
            f = os.path.join(d, mod_name + ext)
            if os.path.isfile(f):
                code = loader_cls(full, f).get_code(full)
                exec(code, module_globals)
                return f
    return None
