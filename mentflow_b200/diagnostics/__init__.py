from .diagnostics import Diagnostic, Histogram, Histogram1D, Histogram2D, Projection
from .histogram import kde_histogram_1d, kde_histogram_2d
