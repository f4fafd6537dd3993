"""TEST INFRASTRUCTURE ONLY: CPU restatement (oracle) of the BDRC page->text hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package -- as the checker, never as the thing measured or shipped.  The product
path (ocr_inference_b200/) never imports it and has no CPU fallback.
"""
