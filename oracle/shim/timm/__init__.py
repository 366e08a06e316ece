"""ORACLE shim (test infrastructure): the handful of timm@a41de1f classes the reference's vision_transformer.py
imports (vision_transformer.py:9-14), restated from SURVEY.md Appendix A.  Forward semantics only; not the product."""
