"""Exception types of the public API (same names as bayesloop/exceptions.py:7-17)."""


class ConfigurationError(Exception):
    """The study set-up is inconsistent (missing model, wrong prior shape, unknown parameter name ...)."""


class PostProcessingError(Exception):
    """A result accessor was called with arguments that do not match the fitted study."""
