// ORACLE shim: placeholder (pcl::fromROSMsg is not restated; the handlers that need it are not compiled)
#pragma once
