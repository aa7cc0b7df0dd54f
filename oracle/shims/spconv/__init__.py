"""TEST INFRASTRUCTURE ONLY — CPU restatement of the spconv 2.x API surface the reference uses
(`spconv-cu117`, requirements.txt:20, version unpinned; call sites models/modules.py:227,242,252,
267,420,444 and models/occupancy_initialization.py:2).  PARITY UNPINNED (see torchsparse shim)."""
