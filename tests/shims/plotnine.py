"""test shim: plotting is outside the hot path; `from plotnine import *` must import and the plot expressions of
mutlib (ggplot(...) + geom_bar(...) + ... ).save(...) must evaluate to nothing"""


class _Layer:
    def __init__(self, *a, **k):
        pass

    def __add__(self, other):
        return self

    __radd__ = __add__

    def save(self, *a, **k):
        pass


__all__ = ["ggplot", "aes", "geom_bar", "theme_bw", "facet_grid", "scale_fill_manual", "labs", "ggtitle", "theme",
           "element_text", "element_blank", "geom_histogram", "geom_point", "geom_line", "scale_x_continuous",
           "scale_y_continuous", "coord_flip", "facet_wrap", "xlab", "ylab"]
for _n in __all__:
    globals()[_n] = type(_n, (_Layer,), {})
