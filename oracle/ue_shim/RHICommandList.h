// SHIM (see CoreMinimal.h in this directory)
#pragma once
#include "RHIResources.h"
