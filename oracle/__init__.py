"""CPU oracle of the MM SAM-Adapter hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain PyTorch-on-CPU (fp32/fp64) restatement of the reference's algorithm, function by function,
each citing the reference file:line it follows. Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import it; the product package
(multimodal-sam-adapter_b200/) never does.

Pinning: oracle/msda.py is checked against the reference's own known-answer vectors
(segmentation/ops/test.py:16-75) and every other function against outputs of the reference's own
modules imported from /root/reference in the build container (tools/make_golden.py ->
tests/golden/*.pt). See DESIGN.md "Oracle".
"""
