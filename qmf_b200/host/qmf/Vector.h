// qmf::Vector lives beside qmf::Matrix (qmf/Matrix.h); this header exists because reference
// callers include <qmf/Vector.h> (qmf/Vector.h:17-46 of the reference).
#pragma once
#include <qmf/Matrix.h>
