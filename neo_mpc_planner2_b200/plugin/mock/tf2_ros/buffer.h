#pragma once
namespace tf2_ros { class Buffer {}; }
