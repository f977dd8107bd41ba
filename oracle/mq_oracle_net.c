#include "mq_oracle.h"
