"""Host-side mirror of the reference's ``code/lib/model`` operator surface.

Same class / method names and argument meaning as the reference; the bodies call the
C-ABI library (``multiply_b200/_lib.py``) — there is no PyTorch or CPU fallback.
"""
