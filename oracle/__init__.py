"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the BTSbot alert-scoring hot path.

Nothing under ``btsbot_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` use it, and only as the checker / the reported CPU baseline.

Pinning status (see DESIGN.md "Oracle"):
* reference-owned glue (`btsbot/architectures.py` metadata branch, fusion head, head
  surgery, `alert_utils` crop/normalise/pad): PINNED -- golden vectors under
  ``tests/golden/`` were produced by executing the reference's own source files in the
  build container (``oracle/make_golden.py``).
* third-party trunk (``timm`` ConvNeXt, ``timm>=0.9.0`` per `pyproject.toml:43`, un-vendored,
  not installable offline): restated from its published architecture and cross-checked
  against torchvision's independent ConvNeXt implementation; the reference has no tests or
  golden vectors for it -> *trunk parity vs timm itself is unpinned*.
"""
