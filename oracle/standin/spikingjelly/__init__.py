"""TEST INFRASTRUCTURE ONLY (oracle stand-in, never imported by the product).

Minimal restatement of the parts of ``spikingjelly==0.0.0.0.14`` that the reference's
model files import (reference requirements.txt:5).  The real package is not installable
in this environment (no network), so the semantics below follow SURVEY.md Appendix A and
the published definitions of the LIF/IF/PLIF neurons and the ATan surrogate.
"""
