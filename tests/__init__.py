"""Hypothesis profiles, as the reference's tests/__init__.py:9-22 defines them."""
import os

from hypothesis import HealthCheck, settings

settings.register_profile("single", max_examples=1, deadline=None)
settings.register_profile("dev", max_examples=25, deadline=None)
settings.register_profile("ci", max_examples=100, deadline=None, suppress_health_check=[HealthCheck.too_slow])
settings.load_profile(os.environ.get("HYPOTHESIS_PROFILE", "dev"))
