// Forwarding header: lets sources that `#include "complex.h"` (reference mandelbrot.h:5)
// pick up the drop-in LPComplex / HPComplex.
#include "mandelbrot.h"
