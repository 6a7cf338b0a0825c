// MetricsManager is declared with the metric classes in qmf/metrics/Metrics.h; reference callers
// include <qmf/metrics/MetricsManager.h> (qmf/metrics/MetricsManager.h:17-72 of the reference).
#pragma once
#include <qmf/metrics/Metrics.h>
