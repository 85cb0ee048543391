"""test shim: plotting is outside the hot path; `from plotnine import *` must just import"""
__all__ = []
