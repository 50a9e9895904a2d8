#include "../cvshim.h"
