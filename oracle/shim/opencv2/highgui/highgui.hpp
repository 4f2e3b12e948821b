// TEST INFRASTRUCTURE ONLY — empty stand-in: FullSystem/CoarseInitializer.cpp includes OpenCV's highgui header but uses nothing from it
#pragma once
