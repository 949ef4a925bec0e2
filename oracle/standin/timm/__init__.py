"""TEST INFRASTRUCTURE ONLY: stand-in for the two ``timm==0.6.13`` symbols the reference uses."""
