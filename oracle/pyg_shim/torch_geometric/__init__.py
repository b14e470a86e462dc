"""TEST INFRASTRUCTURE ONLY (oracle harness) -- not part of the product path.

Minimal restatement of the torch_geometric subset that the reference's
`code/Ob_propagation.py` and `code/transformer_conv.py` import.  torch_geometric is
not installable here (no network) and is unpinned in the reference's
requirements.txt:1-9, so parity at this boundary is "unpinned" (SURVEY.md section 8c):
the semantics below are the ones that are stable across PyG 1.6 .. 2.x for dense
`edge_index` tensors with flow source_to_target.
"""
