// Header name kept from the reference (include/schwarz_base.hpp); the classes live in
// schwz_classes.hpp.
#pragma once
#include "schwz_classes.hpp"
