#pragma once
// real pluginlib registers a factory; the stand-in exports a C factory so the test can dlopen the plugin
#define PLUGINLIB_EXPORT_CLASS(cls, base) extern "C" base * neompc_plugin_create() { return new cls(); }
