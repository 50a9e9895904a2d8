#include <cvshim.h>
