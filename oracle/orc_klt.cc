// ORACLE — TEST INFRASTRUCTURE ONLY. KLT restatement (to be filled).
