// TEST INFRASTRUCTURE: empty stand-in; trackdlo.h names this header but trackdlo.cpp / utils.cpp use nothing from it.
#pragma once
