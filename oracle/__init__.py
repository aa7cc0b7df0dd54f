"""oracle/ — TEST INFRASTRUCTURE ONLY.

CPU restatement of the zhen6618/EPRecon feature-volume hot path, used exclusively as the checker in
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  Nothing under
eprecon_b200/ imports it; the product path fails loudly when the CUDA library is missing.

Layout
  shims/       pure-PyTorch stand-ins for the un-vendored third-party wheels (torchsparse v2.0.0,
               spconv-cu117) -- PARITY UNPINNED, cross-checked against dense ATen ops.
  ref_import.py  imports the UNMODIFIED reference modules from /root/reference over those shims
               (build container only; /root/reference does not exist on the GPU box).
  restate.py   self-contained functional restatement of the hot path and of the panoptic rows that hang
               off it (travels to the GPU box); pinned against reference outputs via tests/golden/
               (three fixtures; the two panoptic ones pin plain-ATen reference code directly).
"""
