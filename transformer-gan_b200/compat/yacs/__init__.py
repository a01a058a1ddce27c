"""Minimal stand-in for the `yacs` package (absent from this image): the subset of ``yacs.config.CfgNode`` that
the reference's ``utils/config_helper.py``, ``utils/config_inference.py``, ``train.py`` and ``generate.py`` use.
Lives under ``compat/`` which goes LAST on ``sys.path``, so a real ``yacs`` install always wins."""
