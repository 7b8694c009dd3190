// cgal_stub: see CGAL/Simple_cartesian.h in this directory
#include <CGAL/Simple_cartesian.h>
