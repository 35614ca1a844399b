"""Import the UNMODIFIED reference modules from /root/reference (this container only).

Used by make_golden.py to generate the committed fixtures; nothing in tests/,
smoke() or bench.py imports this at run time on the GPU box (the reference does
not exist there).

The reference imports two packages that are absent from this image
(`type_enforced`, `torch_geometric`); of those only a no-op decorator, a base
class and two registration decorators are used (fsw_conv.py:4-9, :54-56,
fsw_embedding.py:85, :171).  We install inert stand-ins into sys.modules before
loading the reference by file path, so the reference source itself runs unchanged.
"""
import importlib.util
import os
import sys
import types

REF_DIR = os.environ.get("FSW_REFERENCE_DIR", "/root/reference")


def _install_shims():
    if "type_enforced" not in sys.modules:
        te = types.ModuleType("type_enforced")

        def Enforcer(*args, **kwargs):
            if len(args) == 1 and callable(args[0]) and not kwargs:
                return args[0]
            return lambda f: f

        te.Enforcer = Enforcer
        sys.modules["type_enforced"] = te

    if "torch_geometric" not in sys.modules:
        import torch

        pyg = types.ModuleType("torch_geometric")
        nn = types.ModuleType("torch_geometric.nn")
        utils = types.ModuleType("torch_geometric.utils")
        gg = types.ModuleType("torch_geometric.graphgym")
        reg = types.ModuleType("torch_geometric.graphgym.register")

        class MessagePassing(torch.nn.Module):
            def __init__(self, aggr=None, **kw):
                super().__init__()

        nn.MessagePassing = MessagePassing
        utils.add_self_loops = lambda *a, **k: None
        utils.degree = lambda *a, **k: None
        gg.cfg = None
        reg.register_layer = lambda name: (lambda cls: cls)
        reg.register_pooling = lambda name: (lambda cls: cls)
        gg.register = reg
        pyg.nn, pyg.utils, pyg.graphgym = nn, utils, gg
        sys.modules.update({
            "torch_geometric": pyg,
            "torch_geometric.nn": nn,
            "torch_geometric.utils": utils,
            "torch_geometric.graphgym": gg,
            "torch_geometric.graphgym.register": reg,
        })


def load_reference():
    """Returns (fsw_embedding_module, fsw_conv_module) of the reference."""
    if not os.path.isdir(REF_DIR):
        raise RuntimeError("reference not present at %s" % REF_DIR)
    _install_shims()
    spec = importlib.util.spec_from_file_location("ref_fsw_conv", os.path.join(REF_DIR, "fsw_conv.py"))
    conv = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(conv)
    emb = conv.fsw_embedding
    return emb, conv
