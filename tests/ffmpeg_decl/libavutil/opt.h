#include "avutil_decl.h"
