"""Model families of the message-passing path (mirrors ``deeprank2.neuralnets``)."""
