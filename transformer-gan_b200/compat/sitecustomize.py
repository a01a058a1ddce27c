"""Start-up hooks that let the reference's UNMODIFIED train.py / generate.py import in today's environment
(SURVEY.md section 8b "environment gaps").  Python imports ``sitecustomize`` automatically when this directory is on
PYTHONPATH; nothing here touches the numerical path.

* ``import mem_transformer`` / ``transformer_gan`` / ``discriminator`` / ``utils.proj_adaptive_softmax`` resolve to the
  B200 implementations (a meta-path finder: the script directory would otherwise shadow them).  TGAN_B200_DROPIN=0
  switches this off.

* ``from transformers import AdamW`` (train.py:39-45, utils/classifier.py:5-11): removed from transformers >= 4.x ->
  aliased to ``torch.optim.AdamW`` right after ``transformers`` is imported.
* ``torch.load`` of checkpoints that carry a pickled vocabulary object (train.py:94,620, generate.py:134): torch >= 2.6
  defaults to ``weights_only=True`` -> opt back in through TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD unless the user set it.
* ``--local_rank`` (train.py:122) under torchrun, which only sets LOCAL_RANK: appended to argv when missing.
"""
import importlib.abc
import importlib.util
import os
import sys

os.environ.setdefault("TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD", "1")


class _PatchTransformers(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path, target=None):
        if fullname != "transformers":
            return None
        sys.meta_path.remove(self)
        spec = importlib.util.find_spec("transformers")
        if spec is None or spec.loader is None:
            return None
        loader, orig_exec = spec.loader, spec.loader.exec_module

        def exec_module(module):
            orig_exec(module)
            try:
                # transformers swaps a lazy module object into sys.modules (more than once while it initialises), so
                # the alias is installed on the lazy module CLASS: attribute lookup of "AdamW" that fails the normal way
                # answers torch.optim.AdamW
                live = sys.modules.get("transformers", module)
                cls = type(live)
                orig_getattr = getattr(cls, "__getattr__", None)
                if orig_getattr is not None and not getattr(cls, "_tgan_adamw_alias", False):
                    def __getattr__(self, name, _orig=orig_getattr):
                        if name == "AdamW" and getattr(self, "__name__", "") == "transformers":
                            try:
                                return _orig(self, name)
                            except Exception:  # noqa: BLE001
                                import torch
                                return torch.optim.AdamW
                        return _orig(self, name)
                    cls.__getattr__ = __getattr__
                    cls._tgan_adamw_alias = True
                elif orig_getattr is None and "AdamW" not in vars(live):
                    import torch
                    live.AdamW = torch.optim.AdamW
            except Exception:  # noqa: BLE001
                pass

        loader.exec_module = exec_module
        return spec


sys.meta_path.insert(0, _PatchTransformers())

# The script's own directory (model/) precedes PYTHONPATH on sys.path, so the reference's mem_transformer.py /
# transformer_gan.py would still win a plain path lookup: resolve the hot-path modules to the B200 implementations by name.
_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_DROP_IN = {"mem_transformer": "mem_transformer.py", "transformer_gan": "transformer_gan.py",
            "discriminator": "discriminator.py", "utils.proj_adaptive_softmax": os.path.join("utils", "proj_adaptive_softmax.py")}


class _DropIn(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path, target=None):
        rel = _DROP_IN.get(fullname)
        if rel is None:
            return None
        return importlib.util.spec_from_file_location(fullname, os.path.join(_PKG, rel))


if os.environ.get("TGAN_B200_DROPIN", "1") != "0":
    sys.meta_path.insert(0, _DropIn())
    if _PKG not in sys.path:
        sys.path.append(_PKG)  # tgan_b200 (engine, ctypes binding) and the rest of the package

if "LOCAL_RANK" in os.environ and sys.argv and os.path.basename(sys.argv[0]) == "train.py" and \
        not any(a.startswith("--local_rank") for a in sys.argv):
    sys.argv += ["--local_rank", os.environ["LOCAL_RANK"]]
