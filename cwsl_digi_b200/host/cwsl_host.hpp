// Umbrella header of the host-side mirror of the reference's front-end classes.
#pragma once
#include "CWSL_DIGI_Types.hpp"
#include "WaveFile.hpp"
#include "Decoder.hpp"
#include "DecoderPool.hpp"
#include "IqSource.hpp"
#include "Receiver.hpp"
#include "Instance.hpp"
#include "Decoder_impl.hpp"
#include "Config.hpp"
#include "SlotClock.hpp"
#include "CwslSharedMemory.hpp"
#include "Jt9SharedMemory.hpp"
