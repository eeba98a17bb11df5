// ORACLE shim: placeholder
#pragma once
