"""Legacy import path kept for the integrations (``sdk/decorators_compat.py`` of the reference):
``integrations/huggingface.py:5`` imports ``trace_model_instance`` / ``trace_step`` from here,
and importing it enables the automatic patches once per process, as it always did.
With ``TRACEML_DISABLED=1`` nothing is initialised (no engine, no CUDA requirement)."""
from ..runtime import TraceSessionState, disabled, get_trace_session_state
from .initial import enable_legacy_decorator_auto_init
from .instrumentation import TraceState, trace_model_instance, trace_step, trace_time

if not disabled():
    enable_legacy_decorator_auto_init()

__all__ = ["TraceSessionState", "TraceState", "get_trace_session_state", "trace_step",
           "trace_model_instance", "trace_time"]
