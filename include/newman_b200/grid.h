// Forwarding header: lets sources that `#include "grid.h"` (reference viewer.h:4 via mandelbrot.h:4)
// pick up the drop-in RenderGrid.
#include "mandelbrot.h"
