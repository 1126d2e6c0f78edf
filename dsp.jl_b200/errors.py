"""Exception types mirroring the Julia exceptions the reference throws on this path."""


class ArgumentError(ValueError):
    """Julia ArgumentError (e.g. src/dspbase.jl:28-33, src/periodograms.jl:396)."""


class DomainError(ValueError):
    """Julia DomainError (e.g. src/periodograms.jl:44-45, 397, 565)."""


class DimensionMismatch(ValueError):
    """Julia DimensionMismatch (e.g. src/periodograms.jl:255, 735-737)."""
