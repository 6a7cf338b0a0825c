// MetricsEngine / MetricsConfig are declared in qmf/metrics/Metrics.h; reference callers include
// <qmf/metrics/MetricsEngine.h> (qmf/metrics/MetricsEngine.h:17-135 of the reference).
#pragma once
#include <qmf/metrics/Metrics.h>
