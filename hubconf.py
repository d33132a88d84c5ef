"""torch.hub entry point: ``preprocess, postprocess, model = torch.hub.load(repo, "waternet")``.

Same contract as the reference's ``hubconf.py:37-96`` (return order included);
the work is done by the B200 kernels of ``waternet_b200``.
"""
dependencies = ["torch", "numpy"]

from waternet_b200.hub import arr2ten_noeinops, ten2arr_noeinops, waternet  # noqa: E402,F401
